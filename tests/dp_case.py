"""One rank of the data-parallel parity check (launched by torchrun from tests/test_gpu_round2.py::test_data_parallel_step_on_real_ranks, or by
hand: `python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dp_case.py`).  Every rank prints one RESULT line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from pseudo_speaker_vae_b200.parallel import verify_data_parallel_step

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {}
    for precision, batch in (("fp32", 8192), ("bf16", 16384)):
        r = verify_data_parallel_step(dev, steps=3, global_batch=batch, precision=precision)
        r["rank"] = dist.get_rank()
        out[precision] = r
    # the bar is stated for the fp32 parity mode; bf16 is reported beside it (deterministic sums: only bf16 roundings of d x_hat scaled by
    # 1/B_local vs 1/B differ)
    print("RESULT " + json.dumps(dict(out["fp32"], bf16_param_rel_err=out["bf16"]["param_rel_err"], bf16_grad_rel_err=out["bf16"]["grad_rel_err"],
                                      bf16_ranks_identical=out["bf16"]["ranks_identical"])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
