"""wgrad (contraction over the batch, both operands MN-major) on L2-resident vs HBM-resident operands, and two layouts of the same
operands: columns of a wider [rows][ld] buffer (as in the step: the two encoders share [B][2H] buffers) vs a dense [rows][n] buffer."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_speaker_vae_b200 import _lib as L

def run(rows, out, inn, nsets, reps=20):
    dY = [(torch.randn(rows, out, device="cuda") * 0.1).to(torch.bfloat16) for _ in range(nsets)]
    A = [(torch.randn(rows, inn, device="cuda") * 0.1).to(torch.bfloat16) for _ in range(nsets)]
    g = torch.zeros(out, inn, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f = lambda i: L.check(L.lib().psvae_gemm_probe(dY[i % nsets].data_ptr(), A[i % nsets].data_ptr(), None, g.data_ptr(), None, None, rows, out, inn, 2, st))
    for i in range(4):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        f(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    mb = rows * (out + inn) * 2 / 1e6
    print(f"rows {rows:6d} out {out:4d} in {inn:4d} sets {nsets}: {us:7.1f} us  {mb:6.1f} MB  {mb / us / 1e3:5.2f} TB/s  {2.0 * rows * out * inn / us / 1e6:7.1f} TFLOP/s", flush=True)

for rows, nsets in ((8192, 1), (16384, 1), (32768, 1), (65536, 3), (131072, 3), (262144, 2)):
    run(rows, 512, 512, nsets)
for rows, nsets in ((16384, 1), (65536, 3)):
    run(rows, 1024, 256, nsets)
    run(rows, 256, 512, nsets)
