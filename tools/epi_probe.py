"""Probe of the two hot epilogue forms (psvae_gemm_probe): time per launch for the train step's shapes with the store / mask / column
sums switched off one at a time, operands rotated through buffers larger than L2 (HBM-resident) or kept L2-resident, ring depth sweep."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_speaker_vae_b200 import _lib as L


def run(m, n, k, form, store=True, mask=True, colsum=True, nbuf=4, reps=40, opts=None):
    dev = "cuda"
    for kk, vv in (opts or {}).items():
        L.set_option(kk, vv)
    A = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    W = (torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16) if form == 0 else (torch.randn(k, n, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    O = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    M = [torch.randint(0, 2**31 - 1, (n // 32, m), device=dev, dtype=torch.int32) for _ in range(nbuf)]
    cs = torch.zeros(n, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def call(i):
        j = i % nbuf
        L.check(L.lib().psvae_gemm_probe(A[j].data_ptr(), W.data_ptr(), bias.data_ptr() if form == 0 else None, O[j].data_ptr() if store else None,
                                         M[j].data_ptr() if (mask or form == 1) else None, cs.data_ptr() if (colsum and form == 1) else None,
                                         m, n, k, form, st))
    for i in range(4):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    for kk in (opts or {}):
        L.set_option(kk, 0 if kk not in ("tc_two_cta", "pdl") else 1)
    us = e0.elapsed_time(e1) * 1e3 / reps
    return us, 2.0 * m * n * k / us / 1e6


if __name__ == "__main__":
    B = 65536
    rows = []
    def line(name, *a, **kw):
        us, tf = run(*a, **kw)
        print(f"{name:64s} {us:8.1f} us {tf:8.1f} TFLOP/s", flush=True)
    line("fwd 512x512 HBM full (store+mask)", B, 512, 512, 0)
    line("fwd 512x512 HBM full, no PDL", B, 512, 512, 0, opts={"pdl": 0})
    line("fwd 512x512 HBM store, no mask", B, 512, 512, 0, mask=False)
    line("fwd 512x512 HBM no store, no mask", B, 512, 512, 0, store=False, mask=False)
    line("fwd 512x512 L2-resident (M=16384, 1 buffer) full", 16384, 512, 512, 0, nbuf=1)
    line("fwd 512x512 L2-resident no store no mask", 16384, 512, 512, 0, store=False, mask=False, nbuf=1)
    for s in (2, 3, 4):
        line(f"fwd 512x512 HBM full, ring depth {s}", B, 512, 512, 0, opts={"tc_max_stages": s})
    line("fwd 512x512 HBM full, cta_group::1", B, 512, 512, 0, opts={"tc_two_cta": 0})
    line("fwd 512x512 HBM full, BN=128", B, 512, 512, 0, opts={"tc_force_bn": 128})
    line("fwd K=256 N=1024 HBM full", B, 1024, 256, 0)
    line("fwd K=256 N=1024 HBM no store no mask", B, 1024, 256, 0, store=False, mask=False)
    line("fwd K=64 N=512 HBM full", B, 512, 64, 0)
    line("fwd K=64 N=512 HBM full, no PDL", B, 512, 64, 0, opts={"pdl": 0})
    line("fwd K=64 N=512 HBM no mask", B, 512, 64, 0, mask=False)
    line("fwd K=64 N=512 HBM no store no mask", B, 512, 64, 0, store=False, mask=False)
    line("dgrad 512x512 HBM full (mask+colsum)", B, 512, 512, 1)
    line("dgrad 512x512 HBM no colsum", B, 512, 512, 1, colsum=False)
    line("dgrad 512x512 HBM no store", B, 512, 512, 1, store=False)
    line("dgrad K=64 N=512 HBM full", B, 512, 64, 1)
    line("dgrad K=256 N=512 HBM full", B, 512, 256, 1)
