"""Every GEMM shape of the B = 65,536 train step, timed two ways on the same operands: cuBLAS (torch.matmul, bf16, plain
GEMM without any epilogue) and this repo's tcgen05 engine with its fused epilogue (psvae_gemm_probe).  Operand sets are rotated
(3 x > 126 MB) so that every launch streams its activations from HBM, as inside the step.  Prints one line per shape:
name, FLOPs, cuBLAS us, engine us, the HBM floor (bytes / measured copy bandwidth) and the tensor floor.

    python tools/cublas_compare.py            # on a B200
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_speaker_vae_b200 import _lib as L  # noqa: E402

M = 65536
DEV = "cuda"
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = PEAKS["hbm_gbs"] * 1e9
TF = PEAKS["bf16_tflops"] * 1e12
NSETS = 3


def timed(fn, reps=18):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def bf(*shape):
    return (torch.randn(*shape, device=DEV) * 0.1).to(torch.bfloat16)


def fwd(name, K, N):
    A = [bf(M, K) for _ in range(NSETS)]
    W = bf(N, K)
    bias = torch.zeros(N, device=DEV)
    out = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(NSETS)]
    mask = torch.empty(N // 32 * M, device=DEV, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    t_lib = timed(lambda i: torch.matmul(A[i % NSETS], W.t(), out=out[i % NSETS]))
    t_eng = timed(lambda i: L.check(L.lib().psvae_gemm_probe(A[i % NSETS].data_ptr(), W.data_ptr(), bias.data_ptr(), out[i % NSETS].data_ptr(),
                                                            mask.data_ptr(), None, M, N, K, 0, st)))
    report(name, 2.0 * M * N * K, (M * K + M * N) * 2, t_lib, t_eng)


def dgrad(name, K, N):
    dY = [bf(M, K) for _ in range(NSETS)]
    W = bf(K, N)
    out = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(NSETS)]
    mask = torch.full((N // 32 * M,), -1, device=DEV, dtype=torch.int32)
    cs = torch.zeros(N, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    t_lib = timed(lambda i: torch.matmul(dY[i % NSETS], W, out=out[i % NSETS]))
    t_eng = timed(lambda i: L.check(L.lib().psvae_gemm_probe(dY[i % NSETS].data_ptr(), W.data_ptr(), None, out[i % NSETS].data_ptr(), mask.data_ptr(),
                                                            cs.data_ptr(), M, N, K, 1, st)))
    report(name, 2.0 * M * N * K, (M * K + M * N) * 2 + M * N / 8, t_lib, t_eng)


def wgrad(name, OUT, IN):
    dY = [bf(M, OUT) for _ in range(NSETS)]
    A = [bf(M, IN) for _ in range(NSETS)]
    g16 = torch.empty(OUT, IN, device=DEV, dtype=torch.bfloat16)
    g32 = torch.zeros(OUT, IN, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    t_lib = timed(lambda i: torch.matmul(dY[i % NSETS].t(), A[i % NSETS], out=g16))
    t_eng = timed(lambda i: L.check(L.lib().psvae_gemm_probe(dY[i % NSETS].data_ptr(), A[i % NSETS].data_ptr(), None, g32.data_ptr(), None, None, M, OUT, IN,
                                                            2, st)))
    report(name, 2.0 * M * OUT * IN, (M * OUT + M * IN) * 2, t_lib, t_eng)


def report(name, flops, bytes_, t_lib, t_eng):
    print(f"{name:34s} {flops / 1e9:7.1f} GFLOP  cuBLAS {t_lib:7.1f} us  engine {t_eng:7.1f} us  hbm floor {bytes_ / HBM * 1e6:6.1f} us  tensor floor "
          f"{flops / TF * 1e6:6.1f} us", flush=True)


if __name__ == "__main__":
    print(f"M = {M}, peaks: {PEAKS['hbm_gbs']} GB/s, {PEAKS['bf16_tflops']} TFLOP/s (burst)")
    fwd("fwd enc L0   K=256 N=1024", 256, 1024)
    fwd("fwd hidden   K=512 N=512", 512, 512)
    fwd("fwd dec L0   K=64  N=512", 64, 512)
    fwd("fwd dec last K=512 N=256", 512, 256)
    dgrad("dgrad dec last K=256 N=512", 256, 512)
    dgrad("dgrad hidden   K=512 N=512", 512, 512)
    dgrad("dgrad head     K=64  N=512", 64, 512)
    wgrad("wgrad dec last 256 x 512", 256, 512)
    wgrad("wgrad hidden   512 x 512", 512, 512)
    wgrad("wgrad dec L0   512 x 64", 512, 64)
    wgrad("wgrad head     64 x 512", 64, 512)
    wgrad("wgrad enc L0   1024 x 256", 1024, 256)
