"""Micro-benchmark of the tcgen05 GEMM engine in isolation (psvae_gemm_bf16): TFLOP/s for a few shapes, CTA pairs on/off."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_speaker_vae_b200 import _lib as L


def run(m, n, k, a_mn, b_mn, cg, reps=30, split=1):
    dev = "cuda"
    A = torch.randn(k, m, device=dev).to(torch.bfloat16) if a_mn else torch.randn(m, k, device=dev).to(torch.bfloat16)
    B = torch.randn(k, n, device=dev).to(torch.bfloat16) if b_mn else torch.randn(n, k, device=dev).to(torch.bfloat16)
    C = torch.empty(m, n, device=dev)
    ws = torch.empty(max(1, split) * m * n * 4 + 256, dtype=torch.uint8, device=dev) if split > 1 else None
    L.set_option("tc_two_cta", cg)
    st = torch.cuda.current_stream().cuda_stream
    def call():
        L.check(L.lib().psvae_gemm_bf16(A.data_ptr(), B.data_ptr(), None, C.data_ptr(), m, n, k, int(a_mn), int(b_mn), 0, split, L.ptr(ws), ws.numel() if ws is not None else 0, st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    return us, 2.0 * m * n * k / us / 1e6


if __name__ == "__main__":
    shapes = [
        ("fwd  L2-resident A (16 MB), N=2048", 16384, 2048, 512, 0, 0, 1),
        ("fwd  K=512 full batch", 65536, 512, 512, 0, 0, 1),
        ("fwd  K=256 N=1024 full batch", 65536, 1024, 256, 0, 0, 1),
        ("fwd  big square 8192^3", 8192, 8192, 8192, 0, 0, 1),
        ("dgrad form 65536x512x512", 65536, 512, 512, 0, 1, 1),
        ("wgrad form 512x512x65536 split 18", 512, 512, 65536, 1, 1, 18),
        ("wgrad form 4096x4096x8192", 4096, 4096, 8192, 1, 1, 1),
    ]
    for name, m, n, k, a_mn, b_mn, split in shapes:
        for cg in (0, 1):
            us, tf = run(m, n, k, a_mn, b_mn, cg, split=split)
            print(f"{name:42s} cg={cg}  {us:9.1f} us  {tf:8.1f} TFLOP/s", flush=True)
