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

import os
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


class SymmetricAllReduce:
    """The flat gradient buffer in torch symmetric memory (P2P-mapped on every rank of the node) and its all-reduce through
    ``torch.ops.symm_mem.two_shot_all_reduce_`` (reduce-scatter + all-gather over NVLink peer loads/stores) or ``multimem_all_reduce_``
    (NVLS: the reduction happens inside the NVSwitch).  For the 5 MB gradient these are latency-bound like NCCL's ring, but with one
    kernel and no proxy thread.  ``pick()`` TIMES the candidates on the running ranks (20 calls each) and keeps the fastest -- NCCL if it
    wins or if anything about symmetric memory is unavailable on this node."""

    def __init__(self, numel: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.gname = self.group.group_name
        self.buf = symm.empty(numel, dtype=torch.float32, device=device)
        symm.rendezvous(self.buf, self.gname)
        self.buf.zero_()
        self.kind = "nccl"
        self.timings = {}

    def _call(self, kind: str) -> None:
        if kind == "two_shot":
            torch.ops.symm_mem.two_shot_all_reduce_(self.buf, "sum", self.gname)
        elif kind == "multimem":
            torch.ops.symm_mem.multimem_all_reduce_(self.buf, "sum", self.gname)
        else:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)

    def pick(self, iters: int = 20) -> str:
        dev = self.buf.device
        best = None
        for kind in ("nccl", "two_shot", "multimem"):
            ok = torch.ones(1, device=dev)
            try:
                for _ in range(3):
                    self._call(kind)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier(group=self.group)
                e0.record()
                for _ in range(iters):
                    self._call(kind)
                e1.record()
                torch.cuda.synchronize(dev)
                t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
            except Exception:  # noqa: BLE001 -- this kind is not available here
                ok.zero_()
                t = torch.tensor([float("inf")], device=dev)
            # every rank must take the same decision: a kind counts only if it worked everywhere; its time is the slowest rank's
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            us = float(t) if float(ok) > 0 else float("inf")
            self.timings[kind] = us
            if best is None or us < 0.95 * self.timings[best]:        # a challenger must be 5 % faster than the incumbent (NCCL first)
                best = kind
        self.kind = best
        self.buf.zero_()
        return best

    def all_reduce(self) -> None:
        self._call(self.kind)


class DataParallelTrainer:
    """The per-batch body of the reference's fit loop as three device-side calls:
    fused fwd+bwd  ->  bucketed all-reduce  ->  fused Adam.  Nothing in ``train_step`` synchronises with the host."""

    def __init__(self, module, optimizer=None, bucket_bytes: int = DEFAULT_BUCKET_BYTES, group=None, local_only: bool = False):
        """``local_only``: a trainer that never talks to the other ranks even inside an initialised process group (the single-process
        twin ``verify_data_parallel_step`` compares the data-parallel run with)."""
        self.module = module
        self.hot = module.hot_path
        self.group = group
        self.local_only = bool(local_only)
        self.rank, self.world_size = (0, 1) if self.local_only else world()
        self.bucket_bytes = bucket_bytes
        if optimizer is None:
            optimizer = module.configure_optimizers()["optimizer"]
        self.optimizer = optimizer
        self.optimizer.grad_scale = 1.0 / self.world_size
        if not self.local_only:
            broadcast_parameters(self.hot.arena.ensure(), 0, group)
        self.hot.arena.epoch += 1          # the bf16 operand copy must follow the broadcast values
        self._views = None
        self._views_of = None
        # N > 1: the gradient buffer lives in symmetric memory when one of its all-reduce kernels beats NCCL on these ranks
        # (PSVAE_ALLREDUCE = auto | nccl | two_shot | multimem)
        self.symm: Optional[SymmetricAllReduce] = None
        self.allreduce_kind = "nccl" if self.world_size > 1 else "none"
        want = os.environ.get("PSVAE_ALLREDUCE", "auto")
        if self.world_size > 1 and not self.local_only and want != "nccl" and self.hot.arena.flat.is_cuda:
            sy = None
            ok = torch.ones(1, device=self.hot.arena.flat.device)
            try:
                sy = SymmetricAllReduce(self.hot.arena.numel, self.hot.arena.flat.device, group)
            except Exception:  # noqa: BLE001 -- no symmetric memory on this node / build
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if float(ok) > 0:
                kind = sy.pick() if want == "auto" else want
                sy.kind = kind
                if kind != "nccl":
                    self.symm = sy
                    self.allreduce_kind = kind
                    self.hot.arena._gbuf.append(sy.buf)        # the arena must know the buffer the .grad views live in (FusedAdam.flat_grad)
                self.allreduce_timings = dict(sy.timings)

    def set_shard(self, global_batch: int) -> Tuple[int, int]:
        row0, rows = shard_batch(global_batch, self.rank, self.world_size)
        self.hot.row0 = row0
        return row0, rows

    def train_step(self, x_local: torch.Tensor, y_local=None, eps_local: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimiser step on this rank's shard.  Returns the device tensor of local loss scalars."""
        m, hot = self.module, self.hot
        gflat = self.symm.buf if self.symm is not None else (hot.arena.stage_buffer() if self._views_of is None else self._views_of)
        cons = getattr(m, "consistency_classifier", None)
        losses, gflat, _ = hot.step(x_local, y_local if m.classifier is not None else None, eps_local, kl_weight=m.kl_loss_weight,
                                    clf_weight=m.classifier_loss_weight, use_cos_loss=m.use_cos_loss, compute_grads=True, grads=gflat,
                                    consistency=cons, consistency_y=y_local if cons is not None else None,
                                    consistency_weight=m.consitency_loss_weight)
        if self.symm is not None:
            self.symm.all_reduce()
        elif not self.local_only:
            all_reduce_flat(gflat, self.bucket_bytes, self.group)
        if self._views_of is not gflat:        # bind .grad views once; the same flat buffer is reused every step
            for (p, _), v in zip(hot.arena.entries, hot.arena.grad_views(gflat)):
                p.grad = v if p.requires_grad else None
            self._views_of = gflat
        self.optimizer.step()
        return losses


def verify_data_parallel_step(device, steps: int = 3, global_batch: int = 8192, input_dim: int = 256, latent_dim: int = 64, num_classes: int = 2,
                              precision: str = "fp32", seed: int = 4321, group=None) -> dict:
    """Numerical self-check of the data-parallel path on the ranks that are actually running (what ``DDPStrategy`` guarantees for the
    reference, ps_vae/training.py:78): ``steps`` optimiser steps through ``DataParallelTrainer`` on this rank's shard of a seeded global
    batch with injected eps, then

      * every rank must hold BIT-identical parameters (all-gather of the flat buffer's checksum and of a strided sample), and
      * they must equal a single-process run on the whole global batch from the same initial parameters
        (``||p_dp - p_1|| / ||p_1||`` over the flat parameter buffer; gradients of the last step likewise).

    Runs in the deterministic summation mode (ordered split-K / bias sums) so that the only difference between the two runs is the
    grouping of fp32 partial sums.  Collective: call on every rank.  Returns the measured errors (rank 0 decides what to do with them)."""
    from . import _lib as L
    from .lightning import PseudoSpeakerVAE

    rank, ws = world()
    dev = torch.device(device)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(steps, global_batch, input_dim, generator=g)
    x = x / x.norm(dim=2, keepdim=True)
    y = torch.randint(0, num_classes, (steps, global_batch), generator=g)
    eps = torch.randn(steps, global_batch, latent_dim, generator=g)
    hp = dict(model=dict(input_dim=input_dim, latent_dim=latent_dim), classifier=dict(input_dim=latent_dim, num_classes=num_classes),
              optimizer=dict(lr=1e-3), scheduler=dict(T_max=200), precision=precision)
    prev = L.get_option("deterministic")
    L.set_option("deterministic", 1)
    try:
        torch.manual_seed(seed)                     # same initial weights on every rank and in the single-process twin
        dp = PseudoSpeakerVAE(**hp).to(dev)
        torch.manual_seed(seed)
        one = PseudoSpeakerVAE(**hp).to(dev)
        tr = DataParallelTrainer(dp, group=group)
        row0, rows = tr.set_shard(global_batch)
        solo = DataParallelTrainer(one, local_only=True)           # the twin: same three device-side calls, world of one, no collective
        for s in range(steps):
            xs, ys, es = (t[s].to(dev) for t in (x, y, eps))
            tr.train_step(xs[row0:row0 + rows], ys[row0:row0 + rows], es[row0:row0 + rows])
            solo.hot.row0 = 0
            solo.train_step(xs, ys, es)
        p_dp = dp.hot_path.arena.flat.detach().double()
        p_1 = one.hot_path.arena.flat.detach().double()
        g_dp = dp.hot_path.arena.flat_grad().detach().double() / ws       # the all-reduced SUM over ranks
        g_1 = one.hot_path.arena.flat_grad().detach().double()
        out = dict(world_size=ws, steps=steps, global_batch=global_batch, precision=precision,
                   param_rel_err=float((p_dp - p_1).norm() / p_1.norm()), grad_rel_err=float((g_dp - g_1).norm() / g_1.norm()),
                   param_max_abs=float((p_dp - p_1).abs().max()))
        flat32 = dp.hot_path.arena.flat.detach()
        sig = torch.cat([flat32.view(torch.int32).sum(dtype=torch.int64).reshape(1), flat32.view(torch.int32)[::997].to(torch.int64)])
        if ws > 1:
            sigs = [torch.empty_like(sig) for _ in range(ws)]
            dist.all_gather(sigs, sig, group=group)
            out["ranks_identical"] = bool(all(torch.equal(s, sigs[0]) for s in sigs))
        else:
            out["ranks_identical"] = True
        return out
    finally:
        L.set_option("deterministic", prev)
