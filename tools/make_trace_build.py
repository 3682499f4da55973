"""Builds an INSTRUMENTED copy of the library (build/ab/libpsvae_trace.so) without touching the product sources: the epilogue warps of
the tcgen05 GEMM kernel get cycle counters around (a) the wait for the TMEM read, (b) the wait for the staging buffer's previous TMA
store, (c) staging + fence + column sums + TMA-store issue, (d) the mask / bias fetch before the accumulator wait.  The trace record
grows from 16 to 24 slots per CTA (slots 16..19 = a..d of epilogue warp 0, 20 = tiles that warp processed).

    python tools/make_trace_build.py && PSVAE_B200_LIB=$PWD/build/ab/libpsvae_trace.so python tools/epi_phase_trace.py
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "pseudo_speaker_vae_b200", "csrc")
TMP = os.path.join(ROOT, "build", "ab", "src_trace")       # scratch copy inside the (git-ignored) build directory
OUT = os.path.join(ROOT, "build", "ab", "libpsvae_trace.so")


def sub(s, old, new, count=1):
    assert s.count(old) >= 1, old[:80]
    return s.replace(old, new) if count == 0 else s.replace(old, new, count)


def main():
    shutil.rmtree(TMP, ignore_errors=True)
    shutil.copytree(SRC, os.path.join(TMP, "pseudo_speaker_vae_b200", "csrc"))
    shutil.copytree(os.path.join(ROOT, "include"), os.path.join(TMP, "include"))
    p = os.path.join(TMP, "pseudo_speaker_vae_b200", "csrc", "gemm_tc.cuh")
    s = open(p).read()
    s = sub(s, "(size_t)blockIdx.x * 16", "(size_t)blockIdx.x * 24", 0)
    s = sub(s, "  long long tw0 = 0, tw1 = 0, tw2 = 0;", "  long long tw0 = 0, tw1 = 0, tw2 = 0, tp_ld = 0, tp_buf = 0, tp_st = 0, tp_pre = 0, tp_tiles = 0;")
    # (d) mask / bias prefetch before the accumulator wait
    s = sub(s, "      uint32_t pre[CH];                      // per-block words the functor wants early (EpiActGrad: the ReLU bit masks)",
            "      const long long tq0 = tracing ? clock64() : 0;\n      uint32_t pre[CH];")
    s = sub(s, "      twait(&tfull_bar[acc], acc_phase, 4, tw0);\n      ptx::tc_fence_after();\n      const bool zero_acc = kb0 >= kb1;",
            "      if (tracing) { tp_pre += clock64() - tq0; ++tp_tiles; }\n      twait(&tfull_bar[acc], acc_phase, 4, tw0);\n      ptx::tc_fence_after();\n      const bool zero_acc = kb0 >= kb1;")
    # (a) TMEM read
    s = sub(s, "        ptx::tmem_ld_wait(acc_r);\n        float v[32];",
            "        { const long long tq = tracing ? clock64() : 0; ptx::tmem_ld_wait(acc_r); if (tracing) tp_ld += clock64() - tq; }\n        float v[32];")
    # (b) staging buffer free
    s = sub(s, "          if (opens) {          // the TMA store that last used this staging block must have finished reading it\n            if (lane == 0) ptx::bulk_wait_read0();\n            __syncwarp();\n          }",
            "          if (opens) {\n            const long long tq = tracing ? clock64() : 0;\n            if (lane == 0) ptx::bulk_wait_read0();\n            __syncwarp();\n            if (tracing) tp_buf += clock64() - tq;\n          }\n          const long long tq_st = tracing ? clock64() : 0;")
    # (c) end of the store section: right after the commit block closes
    s = sub(s, "              ptx::bulk_commit();\n            }\n          }\n        }\n      }\n      ptx::tc_fence_before();",
            "              ptx::bulk_commit();\n            }\n          }\n          if (tracing) tp_st += clock64() - tq_st;\n        }\n      }\n      ptx::tc_fence_before();")
    s = sub(s, "    if (warp == 2) { t[7] = (unsigned long long)total; t[8] = (unsigned long long)tw0; t[9] = (unsigned long long)tw1; }",
            "    if (warp == 2) { t[7] = (unsigned long long)total; t[8] = (unsigned long long)tw0; t[9] = (unsigned long long)tw1; t[16] = (unsigned long long)tp_ld; "
            "t[17] = (unsigned long long)tp_buf; t[18] = (unsigned long long)tp_st; t[19] = (unsigned long long)tp_pre; t[20] = (unsigned long long)tp_tiles; }")
    open(p, "w").write(s)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", OUT,
           os.path.join(TMP, "pseudo_speaker_vae_b200", "csrc", "psvae_b200.cu")]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit(1)
    print("built", OUT)


if __name__ == "__main__":
    main()
