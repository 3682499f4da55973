"""Unconditional / conditional sampling throughput with the chained decoder kernel on and off (option "decode_chain").
    python tools/sample_bench.py [N]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseudo_speaker_vae_b200 as P
from pseudo_speaker_vae_b200 import _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
torch.manual_seed(0)
m = P.PseudoSpeakerVAE(model=dict(input_dim=256, latent_dim=64), classifier=dict(input_dim=64, num_classes=2), optimizer=dict(lr=1e-3),
                       scheduler=dict(T_max=10), precision="bf16").to("cuda")
out = torch.empty(N, 256, device="cuda")
for chain in (0, 1, 0, 1):
    L.set_option("decode_chain", chain)
    for n in (N, 65536 * 4):
        o = out[:n]
        P.sample_on_device(m, n, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if n == N else 20
        e0.record()
        for _ in range(reps):
            P.sample_on_device(m, n, out=o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"decode_chain={chain} N={n}: {ms:.3f} ms  {n / ms / 1e6:.3f} G samples/s  tensor {n / ms / 1e6 * 851968 / 1e3 / 1599.4:.3f} of burst", flush=True)
z = torch.randn(1 << 20, 64, device="cuda")
for chain in (0, 1):
    L.set_option("decode_chain", chain)
    m.hot_path.decode(z, out=out[:1 << 20])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        m.hot_path.decode(z, out=out[:1 << 20])
    e1.record()
    torch.cuda.synchronize()
    print(f"decode(z given) chain={chain}: {e0.elapsed_time(e1) / 10:.3f} ms per 1 Mi rows", flush=True)
