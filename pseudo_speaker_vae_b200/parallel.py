"""Data-parallel training of the fused step: one process per GPU, replicated parameters, batch rows sharded
across ranks, ONE bucketed all-reduce of the flat gradient buffer per step (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  This is what ``Trainer(strategy=DDPStrategy(...))`` does for the reference
(ps_vae/training.py:78), minus its per-parameter bookkeeping:

  C1  gradient all-reduce   -> ``all_reduce_flat``: SUM over ranks of the arena's flat gradient buffer in
                               ``bucket_bytes`` slices (25 MiB like DDP: 5.13 MB of gradients = one bucket),
                               averaging folded into Adam's ``grad_scale = 1/world``;
  C2  per-metric scalar all-reduces (``sync_dist=True``, 3-7 per step) -> the 16 loss slots ride in ONE small
                               all-reduce (``reduce_losses``), only when somebody asks for them;
  C3  initial parameter broadcast -> ``broadcast_parameters`` (one broadcast of the flat buffer);
  C4  DistributedSampler    -> ``shard_batch``: rank r owns rows [r*B/W, (r+1)*B/W) and tells the kernels its
                               global row offset, so the Philox eps stream does not depend on the world size.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

DEFAULT_BUCKET_BYTES = 25 * 1024 * 1024


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(row0, rows) of this rank's contiguous shard; the batch must split evenly (DDP semantics need equal shards
    for the mean of means to equal the global mean)."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} does not split evenly over {world_size} ranks")
    rows = global_batch // world_size
    return rank * rows, rows


def bucket_slices(numel: int, bucket_bytes: int = DEFAULT_BUCKET_BYTES, elem_bytes: int = 4):
    per = max(1, bucket_bytes // elem_bytes)
    return [(a, min(numel, a + per)) for a in range(0, numel, per)]


def all_reduce_flat(flat: torch.Tensor, bucket_bytes: int = DEFAULT_BUCKET_BYTES, group=None, async_op: bool = False):
    """SUM all-reduce of a flat buffer in buckets.  Returns the work handles when ``async_op``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    works = []
    for a, b in bucket_slices(flat.numel(), bucket_bytes, flat.element_size()):
        w = dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def broadcast_parameters(flat: torch.Tensor, src: int = 0, group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)


def reduce_losses(losses: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over ranks of the packed loss scalars (equal shards): one 64-byte all-reduce instead of 3-7."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = losses.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out / dist.get_world_size(group)
    return losses


class DataParallelTrainer:
    """The per-batch body of the reference's fit loop as three device-side calls:
    fused fwd+bwd  ->  bucketed all-reduce  ->  fused Adam.  Nothing in ``train_step`` synchronises with the host."""

    def __init__(self, module, optimizer=None, bucket_bytes: int = DEFAULT_BUCKET_BYTES, group=None):
        self.module = module
        self.hot = module.hot_path
        self.group = group
        self.rank, self.world_size = world()
        self.bucket_bytes = bucket_bytes
        if optimizer is None:
            optimizer = module.configure_optimizers()["optimizer"]
        self.optimizer = optimizer
        self.optimizer.grad_scale = 1.0 / self.world_size
        broadcast_parameters(self.hot.arena.ensure(), 0, group)
        self.hot.arena.epoch += 1          # the bf16 operand copy must follow the broadcast values
        self._views = None
        self._views_of = None

    def set_shard(self, global_batch: int) -> Tuple[int, int]:
        row0, rows = shard_batch(global_batch, self.rank, self.world_size)
        self.hot.row0 = row0
        return row0, rows

    def train_step(self, x_local: torch.Tensor, y_local=None, eps_local: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimiser step on this rank's shard.  Returns the device tensor of local loss scalars."""
        m, hot = self.module, self.hot
        gflat = hot.arena.stage_buffer() if self._views_of is None else self._views_of
        cons = getattr(m, "consistency_classifier", None)
        losses, gflat, _ = hot.step(x_local, y_local if m.classifier is not None else None, eps_local, kl_weight=m.kl_loss_weight,
                                    clf_weight=m.classifier_loss_weight, use_cos_loss=m.use_cos_loss, compute_grads=True, grads=gflat,
                                    consistency=cons, consistency_y=y_local if cons is not None else None,
                                    consistency_weight=m.consitency_loss_weight)
        all_reduce_flat(gflat, self.bucket_bytes, self.group)
        if self._views_of is not gflat:        # bind .grad views once; the same flat buffer is reused every step
            for (p, _), v in zip(hot.arena.entries, hot.arena.grad_views(gflat)):
                p.grad = v if p.requires_grad else None
            self._views_of = gflat
        self.optimizer.step()
        return losses
