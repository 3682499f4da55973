"""Where the GEMM kernel's warps wait: runs one probe launch with the cycle counters on (option tc_trace_ptr) and prints, averaged over
CTAs, the cycles each role spent in its barrier waits.  python tools/tc_trace.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseudo_speaker_vae_b200 import _lib as L


def trace(m, n, k, form, store=True, mask=True, colsum=True, opts=None, label=""):
    dev = "cuda"
    for kk, vv in (opts or {}).items():
        L.set_option(kk, vv)
    nb = 3
    A = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nb)]
    W = (torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16) if form == 0 else (torch.randn(k, n, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    O = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    M = [torch.randint(0, 2**31 - 1, (n // 32, m), device=dev, dtype=torch.int32) for _ in range(nb)]
    cs = torch.zeros(n, device=dev)
    tr = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    tr2 = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def call(j):
        L.check(L.lib().psvae_gemm_probe(A[j].data_ptr(), W.data_ptr(), bias.data_ptr() if form == 0 else None, O[j].data_ptr() if store else None,
                                         M[j].data_ptr() if (mask or form == 1) else None, cs.data_ptr() if (colsum and form == 1) else None,
                                         m, n, k, form, st))
    for j in range(nb):
        call(j)
    torch.cuda.synchronize()
    L.set_option("tc_trace_ptr", tr.data_ptr())
    call(0)
    L.set_option("tc_trace_ptr", tr2.data_ptr())
    call(1)
    torch.cuda.synchronize()
    L.set_option("tc_trace_ptr", 0)
    for kk in (opts or {}):
        L.set_option(kk, 0 if kk not in ("tc_two_cta", "pdl") else 1)
    t = tr.view(148, 16).double().cpu()
    t2 = tr2.view(148, 16).double().cpu()
    ent, beg, epi_end, ext = t[:, 12], t[:, 13], t[:, 14], t[0::2, 15]
    k0 = ent.min().item()
    print(f"   timeline (us from first CTA entry): entry spread {(ent.max().item()-k0)/1e3:6.2f} | role loops start {(beg.min().item()-k0)/1e3:6.2f}..{(beg.max().item()-k0)/1e3:6.2f} | "
          f"epilogue done {(epi_end.min().item()-k0)/1e3:6.2f}..{(epi_end.max().item()-k0)/1e3:6.2f} | exit max {(ext.max().item()-k0)/1e3:6.2f} | "
          f"next kernel first entry {(t2[:,12].min().item()-k0)/1e3:6.2f}")
    lead = t[0::2]      # leaders of the pairs (MMA warps run there)
    f = lambda x: f"{x.mean().item():9.0f}"
    print(f"{label:46s} MMA-total {f(lead[:,3])} epi-total {f(t[:,7])} producer-total {f(t[:,0])} | producer wait-empty {f(t[:,1])} | MMA(leader) wait-full {f(lead[:,4])} wait-tempty {f(lead[:,5])} "
          f"wait-bfull {f(lead[:,6])} | epi w0 wait-tfull {f(t[:,8])} aux {f(t[:,9])} | epi w3 wait-tfull {f(t[:,11])}", flush=True)


def trace_wgrad(m, n, k, label=""):
    dev = "cuda"
    nb = 3
    A = [torch.randn(m, n, device=dev).to(torch.bfloat16) for _ in range(nb)]
    W = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nb)]
    G = torch.zeros(n, k, device=dev)
    tr = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    tr2 = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def call(j):
        L.check(L.lib().psvae_gemm_probe(A[j].data_ptr(), W[j].data_ptr(), None, G.data_ptr(), None, None, m, n, k, 2, st))
    for j in range(nb):
        call(j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(30):
        call(j % nb)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 30
    L.set_option("tc_trace_ptr", tr.data_ptr())
    call(0)
    L.set_option("tc_trace_ptr", tr2.data_ptr())
    call(1)
    torch.cuda.synchronize()
    L.set_option("tc_trace_ptr", 0)
    t = tr.view(148, 16).double().cpu()
    t2 = tr2.view(148, 16).double().cpu()
    act = t[:, 12] > 0
    t = t[act]
    ent, beg, epi_end = t[:, 12], t[:, 13], t[:, 14]
    k0 = ent.min().item()
    lead = t[0::2]
    f = lambda x: f"{x.mean().item():9.0f}"
    print(f"   wgrad {us:6.1f} us/launch ({2.0*m*n*k/us/1e6:6.1f} TFLOP/s, {2.0*m*(n+k)/us/1e6:5.2f} TB/s operand bytes) CTAs {int(act.sum())} | loops start {(beg.min().item()-k0)/1e3:5.2f}..{(beg.max().item()-k0)/1e3:5.2f} "
          f"| epilogue done {(epi_end.min().item()-k0)/1e3:6.2f}..{(epi_end.max().item()-k0)/1e3:6.2f} | next entry {(t2[:,12][t2[:,12]>0].min().item()-k0)/1e3:6.2f}")
    print(f"{label:46s} MMA-total {f(lead[:,3])} epi-total {f(t[:,7])} producer-total {f(t[:,0])} | producer wait-empty {f(t[:,1])} | MMA wait-full {f(lead[:,4])} "
          f"wait-tempty {f(lead[:,5])} | epi w0 wait-tfull {f(t[:,8])}", flush=True)


if __name__ == "__main__":
    trace_wgrad(65536, 512, 512, label="wgrad 512x512 (K = 65536 rows)")
    trace_wgrad(65536, 1024, 256, label="wgrad 1024x256")
    trace_wgrad(65536, 256, 512, label="wgrad 256x512")
    trace_wgrad(65536, 64, 512, label="wgrad 64x512")
    B = 65536
    trace(B, 512, 512, 0, label="fwd 512x512 full")
    trace(B, 512, 512, 0, label="fwd 512x512 full no PDL", opts={"pdl": 0})
    trace(B, 512, 512, 0, store=False, mask=False, label="fwd 512x512 no store no mask")
    trace(B, 1024, 256, 0, label="fwd K=256 N=1024 full")
    trace(B, 512, 64, 0, label="fwd K=64 N=512 full")
    trace(B, 512, 64, 0, store=False, mask=False, label="fwd K=64 N=512 no store no mask")
    trace(B, 512, 512, 1, label="dgrad 512x512 full")
    trace(16384, 512, 512, 0, label="fwd 512x512 M=16384")
