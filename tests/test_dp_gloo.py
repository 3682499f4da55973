"""world_size-2 gloo tests of the data-parallel plumbing (pseudo_speaker_vae_b200/parallel.py) on the CPU:
the bucketed flat all-reduce, the parameter broadcast, the packed loss reduce -- and that sharding rows over ranks and
averaging the per-shard gradients reproduces the full-batch gradient (the property the GPU step relies on).  The
per-shard gradients come from the numpy oracle here (there is no GPU in this container); the collective code is the product's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ps_vae_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pseudo_speaker_vae_b200 as P
        from pseudo_speaker_vae_b200 import parallel

        torch.manual_seed(100 + rank)          # ranks start from DIFFERENT weights: the broadcast must fix that
        m = P.PseudoSpeakerVAE(model=dict(input_dim=64, latent_dim=8, hidden_dim=64), classifier=dict(input_dim=8, num_classes=3),
                               optimizer=dict(lr=1e-3), scheduler=dict(T_max=10))
        hot = m.hot_path
        flat = hot.arena.ensure()
        parallel.broadcast_parameters(flat, 0)
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        assert hot.arena.attached()            # the views the modules hold saw the broadcast

        params = {k: v.detach().double().numpy() for k, v in m.state_dict().items()}
        B = 48
        x, y, eps = O.synth_batch(B, 64, 8, 3, seed=5, dtype=np.float64)
        row0, rows = parallel.shard_batch(B, rank, world)
        sl = slice(row0, row0 + rows)
        scal, _, grads = O.train_loss_and_grads(params, x[sl], y[sl], eps[sl])
        names = {id(p): k for k, p in m.named_parameters()}
        g = torch.zeros(hot.arena.numel, dtype=torch.float32)
        for p, off in hot.arena.entries:
            g[off:off + p.numel()] = torch.from_numpy(grads[names[id(p)]].reshape(-1)).float()
        for bucket in (25 * 1024 * 1024, 4096):          # one bucket / many buckets
            gg = g.clone()
            parallel.all_reduce_flat(gg, bucket)
            gg /= world
            _, _, full = O.train_loss_and_grads(params, x, y, eps)
            for p, off in hot.arena.entries:
                ref = full[names[id(p)]].reshape(-1)
                got = gg[off:off + p.numel()].double().numpy()
                assert np.sqrt(((got - ref) ** 2).sum()) <= 1e-6 * max(np.sqrt((ref ** 2).sum()), 1e-12), names[id(p)]
        works = parallel.all_reduce_flat(g.clone(), 4096, async_op=True)
        assert len(works) == len(parallel.bucket_slices(g.numel(), 4096))
        for w in works:
            w.wait()
        losses = torch.zeros(16)
        losses[0] = float(scal["loss"])
        red = parallel.reduce_losses(losses)
        s_full, _, _ = O.train_loss_and_grads(params, x, y, eps, compute_grads=False)
        assert abs(float(red[0]) - float(s_full["loss"])) < 1e-6          # equal shards: mean of means == global mean
        # sampling shards cover the index range exactly once
        spans = [P.shard_rows(1001, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == 1001
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_data_parallel_plumbing_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))
