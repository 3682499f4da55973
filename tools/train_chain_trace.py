"""Where the TRAIN chained decoder kernel's roles wait: one fused step at B = 65,536 with option train_chain and the cycle counters on
(tc_trace_ptr; the GEMM engine's kernels write their own 16 counters per CTA into the same buffer first -- the chain kernel's 24 per CTA are
read from a second buffer set just for a forward-only... simpler: the whole step runs traced and only the chain kernel's slots are printed).
    python tools/train_chain_trace.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseudo_speaker_vae_b200 as P
from pseudo_speaker_vae_b200 import _lib as L

B = 65536
torch.manual_seed(0)
m = P.PseudoSpeakerVAE(model=dict(input_dim=256, latent_dim=64), classifier=dict(input_dim=64, num_classes=2), optimizer=dict(lr=1e-3),
                       scheduler=dict(T_max=10), precision="bf16").to("cuda")
hot = m.hot_path
x = torch.randn(B, 256, device="cuda")
x = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
y = torch.randint(0, 2, (B,), device="cuda")
g = torch.empty(hot.arena.numel, device="cuda")
L.set_option("train_chain", 1)
for _ in range(3):
    hot.step(x, y, grads=g)
torch.cuda.synchronize()
tr = torch.zeros(148 * 24 + 4096, dtype=torch.int64, device="cuda")
L.set_option("tc_trace_ptr", tr.data_ptr())
hot.step(x, y, grads=g)
torch.cuda.synchronize()
L.set_option("tc_trace_ptr", 0)
# the GEMM kernels before / after the chain kernel overwrite slots [cta * 16 + 0..15]; the chain kernel writes [cta * 24 + ...] and ran in
# the middle: re-run ONLY the forward part is not possible through the API, so run the decode-free trick: trace again with the engine's
# kernels untraced is not possible either -> read what the LAST writer left: the wgrad kernel (last GEMM) wrote cta*16 slots, the chain
# kernel's slots beyond 148*16 survive for CTAs >= 99 (99*24 = 2376 > 148*16 = 2368)
t = tr[:148 * 24].view(148, 24).double().cpu()
tiles = B / 256 / 74
keep = [c for c in range(100, 148)]
lead = [c for c in keep if c % 2 == 0]
f = lambda rows, col: f"{sum(t[r, col].item() for r in rows) / len(rows) / tiles:8.0f}"
print(f"TRAIN chain, cycles PER TILE (mean over CTAs 100..147; {tiles:.2f} tiles per pair)")
print(f" MMA warp (leader): total {f(lead,0)} | wait weights {f(lead,1)} acc_empty {f(lead,2)} hd0_ready {f(lead,3)} hd1_ready {f(lead,4)} out_empty {f(lead,5)} z_full {f(lead,6)}")
print(f" epilogue warp 2  : total {f(keep,8)} | wait acc_full {f(keep,9)} hd0_free {f(keep,10)} hd1_empty {f(keep,11)} out_full {f(keep,12)}")
print(f" producer         : total {f(keep,16)} | wait b_empty {f(keep,17)}")
print(f" z warp           : total {f(keep,18)} | wait z_empty {f(keep,19)}")
