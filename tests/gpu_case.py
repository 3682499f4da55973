"""Stand-alone GPU cases run in their own process (a tcgen05 protocol bug traps the whole CUDA context, so the
tensor-core GEMM variants are isolated from the rest of the suite).  Prints one JSON line.

    python tests/gpu_case.py gemm --m 300 --n 192 --k 520 --a_mn 0 --b_mn 1 --bn 64 --split 1
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--m", type=int, default=256)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--a_mn", type=int, default=0)
    ap.add_argument("--b_mn", type=int, default=0)
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--split", type=int, default=1)
    ap.add_argument("--relu", type=int, default=0)
    ap.add_argument("--bias", type=int, default=0)
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--cg", type=int, default=0, help="1: CTA pairs (tcgen05 cta_group::2)")
    a = ap.parse_args()
    import torch

    from pseudo_speaker_vae_b200 import _lib as L
    from tests.gpu_util import gemm_bf16

    out = dict(case=a.case, args=vars(a))
    if a.case == "gemm":
        torch.manual_seed(a.m * 7 + a.n * 3 + a.k)
        A = torch.randn(a.m, a.k, device="cuda")
        B = torch.randn(a.n, a.k, device="cuda")
        bias = torch.randn(a.n, device="cuda") if a.bias else None
        L.set_option("tc_force_bn", a.bn)
        L.set_option("tc_grid_limit", a.grid)
        L.set_option("tc_two_cta", a.cg)
        c, ref = gemm_bf16(A, B, bias, bool(a.a_mn), bool(a.b_mn), bool(a.relu), a.split)
        torch.cuda.synchronize()
        err = ((c.double() - ref).norm() / ref.norm()).item()
        maxabs = (c.double() - ref).abs().max().item()
        out.update(rel_err=err, max_abs=maxabs, finite=bool(torch.isfinite(c).all().item()))
    else:
        raise SystemExit(f"unknown case {a.case}")
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
