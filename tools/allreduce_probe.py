"""All-reduce of the flat gradient buffer (1,281,600 fp32 = 5.13 MB): NCCL against torch's symmetric-memory kernels (P2P one-shot / two-shot, NVLS multimem)
on the ranks that are running.   torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, ws = dist.get_rank(), dist.get_world_size()
n = 1281600
gname = dist.group.WORLD.group_name
def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
g = torch.randn(n, device=dev)
res = {"nccl": timeit(lambda: dist.all_reduce(g))}
try:
    t = symm.empty(n, dtype=torch.float32, device=dev)
    symm.rendezvous(t, gname)
    t.copy_(torch.randn(n, device=dev))
    ref = t.clone(); dist.all_reduce(ref)
    for name, fn in (("two_shot", lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname)),
                     ("one_shot", lambda: torch.ops.symm_mem.one_shot_all_reduce(t, "sum", gname)),
                     ("multimem", lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname))):
        try:
            if name == "two_shot":
                chk = t.clone(); 
            res[name] = timeit(fn)
        except Exception as e:
            res[name] = "failed: " + repr(e)[:120]
    # correctness of two_shot on fresh data
    t.copy_(torch.arange(n, device=dev, dtype=torch.float32) * (rank + 1) * 1e-3)
    want = t.clone(); dist.all_reduce(want)
    torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname); torch.cuda.synchronize()
    res["two_shot_max_err"] = float((t - want).abs().max())
except Exception as e:
    res["symm"] = "failed: " + repr(e)[:200]
if rank == 0:
    print("allreduce 5.13 MB fp32, world", ws, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}, "(us)", flush=True)
dist.barrier(); dist.destroy_process_group()
