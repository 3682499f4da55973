"""Where the chained decoder kernel's roles wait (option tc_trace_ptr: 24 cycle counters per CTA).  python tools/chain_trace.py [N]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseudo_speaker_vae_b200 as P
from pseudo_speaker_vae_b200 import _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 74 * 40
torch.manual_seed(0)
m = P.PseudoSpeakerVAE(model=dict(input_dim=256, latent_dim=64), classifier=dict(input_dim=64, num_classes=2), optimizer=dict(lr=1e-3),
                       scheduler=dict(T_max=10), precision="bf16").to("cuda")
out = torch.empty(N, 256, device="cuda")
L.set_option("decode_chain", 1)
for _ in range(3):
    P.sample_on_device(m, N, out=out)
torch.cuda.synchronize()
tr = torch.zeros(148 * 24, dtype=torch.int64, device="cuda")
L.set_option("tc_trace_ptr", tr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
P.sample_on_device(m, N, out=out)
e1.record()
torch.cuda.synchronize()
L.set_option("tc_trace_ptr", 0)
t = tr.view(148, 24).double().cpu()
lead = t[0::2]
tiles = (N + 255) // 256 / 74
f = lambda x: f"{x.mean().item() / tiles:8.0f}"
print(f"N={N}: {e0.elapsed_time(e1):.3f} ms, {tiles:.1f} tiles per pair; cycles PER TILE (mean over CTAs)")
print(f" MMA warp (leader): total {f(lead[:,0])} | wait weights {f(lead[:,1])} acc_empty {f(lead[:,2])} hd0_ready {f(lead[:,3])} hd1_ready {f(lead[:,4])} out_empty {f(lead[:,5])} z_full {f(lead[:,6])}")
print(f" epilogue warp 2  : total {f(t[:,8])} | wait acc_full {f(t[:,9])} hd0_free {f(t[:,10])} hd1_empty {f(t[:,11])} out_full {f(t[:,12])}")
print(f" producer         : total {f(t[:,16])} | wait b_empty {f(t[:,17])}")
print(f" z warp           : total {f(t[:,18])} | wait z_empty {f(t[:,19])}")
