"""Where an epilogue warp of the tcgen05 GEMM kernel spends its cycles, per tile, on the thin (K = 64) and square (K = 512) layer shapes.
Needs the instrumented build:  python tools/make_trace_build.py && PSVAE_B200_LIB=$PWD/build/ab/libpsvae_trace.so python tools/epi_phase_trace.py
Columns (cycles per tile of epilogue warp 0, averaged over CTAs): total, wait for the accumulator (tfull), mask/bias prefetch issue, wait for the
TMEM read, wait for the staging buffer's previous TMA store, staging + fence + column sums + TMA-store issue."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_speaker_vae_b200 import _lib as L

SLOTS = 24


def run(m, n, k, form, store=True, mask=True, colsum=True, label=""):
    dev = "cuda"
    nb = 3
    A = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nb)]
    W = (torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16) if form == 0 else (torch.randn(k, n, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    O = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    M = [torch.randint(0, 2**31 - 1, (n // 32, m), device=dev, dtype=torch.int32) for _ in range(nb)]
    cs = torch.zeros(n, device=dev)
    tr = torch.zeros(148 * SLOTS, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call(j):
        L.check(L.lib().psvae_gemm_probe(A[j].data_ptr(), W.data_ptr(), bias.data_ptr() if form == 0 else None, O[j].data_ptr() if store else None,
                                         M[j].data_ptr() if (mask or form == 1) else None, cs.data_ptr() if (colsum and form == 1) else None, m, n, k, form, st))
    for j in range(nb):
        call(j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(12):
        call(j % nb)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 12
    L.set_option("tc_trace_ptr", tr.data_ptr())
    call(0)
    torch.cuda.synchronize()
    L.set_option("tc_trace_ptr", 0)
    t = tr.view(148, SLOTS).double().cpu()
    tiles = t[:, 20].clamp(min=1)
    per = lambda i: (t[:, i] / tiles).mean().item()
    print(f"{label:44s} {us:6.1f} us | tiles/CTA {tiles.mean().item():4.2f} | per tile: total {per(7):7.0f}  tfull {per(8):6.0f}  pre {per(19):6.0f}  tmem-ld {per(16):6.0f}  "
          f"buf-wait {per(17):6.0f}  stage+store {per(18):6.0f}  other {per(7) - per(8) - per(19) - per(16) - per(17) - per(18):6.0f}", flush=True)


if __name__ == "__main__":
    B = 65536
    run(B, 512, 64, 0, label="fwd K=64 N=512 full (bias, relu, mask, store)")
    run(B, 512, 64, 0, mask=False, label="fwd K=64 N=512 no mask")
    run(B, 512, 64, 0, store=False, mask=False, label="fwd K=64 N=512 no store no mask")
    run(B, 512, 64, 1, label="dgrad K=64 N=512 full (mask, colsum, store)")
    run(B, 512, 64, 1, colsum=False, label="dgrad K=64 N=512 no colsum")
    run(B, 512, 512, 0, label="fwd K=512 N=512 full")
    run(B, 512, 512, 1, label="dgrad K=512 N=512 full")
