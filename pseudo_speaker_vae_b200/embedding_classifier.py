"""``EmbeddingClassifier`` -- drop-in for ps_vae/embedding_classifier/embedding_classifier.py:6-62 as far as the VAE hot
path needs it: the frozen ``consistency_classifier`` of ``PseudoSpeakerVAE`` (ps_vae/lightning.py:44-52, 100-108).

Same constructor, attributes and state-dict keys (``fc1 / fc2 / fc3``), so a Lightning checkpoint of the reference's
classifier loads unchanged (``load_from_checkpoint``).  ``forward`` runs ``fc3(relu(fc2(relu(fc1(x)))))`` through the
CUDA library (``psvae_consistency_forward``, fp32); inside the fused train step the same weights are read by
``psvae_train_fwd_bwd_consistency`` (engine.HotPath.step).  There is no CPU path.

The classifier's own trainer (embedding_classifier.py:64-100) is here too: ``training_step`` /
``validation_step`` make one library call (``psvae_embedding_classifier_step``: logits, CrossEntropyLoss, accuracy and the six
gradient tensors, fp32) and log ``train_acc / train_loss / val_acc / val_loss`` as the reference does; the returned loss hands the
already-computed gradients to autograd on ``backward()``, and ``configure_optimizers`` is the reference's ``torch.optim.Adam``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from ._compat import _Accuracy, _Base


class EmbeddingClassifier(_Base):
    def __init__(self, input_dim: int, num_classes: int, hidden_dim: int = 128, optimizer_cfg: dict = {}) -> None:  # noqa: B006 (reference signature)
        super().__init__()
        self.save_hyperparameters()
        self.input_dim = int(input_dim)
        self.num_classes = int(num_classes)
        self.hidden_dim = int(hidden_dim)
        self.optimizer_cfg = optimizer_cfg
        self.fc1 = nn.Linear(self.input_dim, self.hidden_dim)
        self.fc2 = nn.Linear(self.hidden_dim, self.hidden_dim)
        self.fc3 = nn.Linear(self.hidden_dim, self.num_classes)
        self.relu = nn.ReLU()
        self.softmax = nn.Softmax(dim=1)
        self.loss_fn = nn.CrossEntropyLoss()
        self.accuracy = _Accuracy(task="multiclass", num_classes=self.num_classes)
        object.__setattr__(self, "_flat_cache", None)

    # ---- flat fp32 copy of the (frozen) weights in the layout psvae_consistency_desc_init defines ------------------
    def consistency_desc(self) -> L.ConsistencyDesc:
        d = getattr(self, "_cdesc", None)
        if d is None:
            d = L.make_consistency_desc(self.input_dim, self.hidden_dim, self.num_classes)
            object.__setattr__(self, "_cdesc", d)
        return d

    def flat_params(self, device: torch.device) -> Tuple[L.ConsistencyDesc, torch.Tensor]:
        """(desc, flat device buffer); rebuilt when a weight changed (``load_state_dict``, ``.to()``) since the last call."""
        d = self.consistency_desc()
        tensors = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.fc3.weight, self.fc3.bias]
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        cache = self._flat_cache
        if cache is None or cache[0] != key:
            flat = torch.zeros(int(d.total_numel), dtype=torch.float32, device=device)
            offs = [d.w[0], d.b[0], d.w[1], d.b[1], d.w[2], d.b[2]]
            with torch.no_grad():
                for t, off in zip(tensors, offs):
                    flat[off:off + t.numel()].copy_(t.detach().reshape(-1).to(device=device, dtype=torch.float32))
            cache = (key, flat)
            object.__setattr__(self, "_flat_cache", cache)
        return d, cache[1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """logits [batch, num_classes] (embedding_classifier.py:50-62)."""
        dev = self.fc1.weight.device
        if dev.type != "cuda":
            raise RuntimeError(f"pseudo_speaker_vae_b200 runs on a B200 only (module is on {dev}); there is no CPU fallback")
        if x.dim() != 2 or x.shape[1] != self.input_dim:
            raise ValueError(f"x must be [batch, {self.input_dim}], got {tuple(x.shape)}")
        if x.device != dev:
            raise ValueError(f"x is on {x.device}, the model on {dev}")
        x = x.detach().to(torch.float32).contiguous()
        rows = x.shape[0]
        logits = torch.empty(rows, self.num_classes, dtype=torch.float32, device=dev)
        if rows == 0:
            return logits
        d, flat = self.flat_params(dev)
        need = int(L.lib().psvae_consistency_workspace_bytes(C.byref(d), rows, L.MODE_FORWARD))
        if need < 0:
            raise ValueError(L.last_error())
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        rc = L.lib().psvae_consistency_forward(C.byref(d), flat.data_ptr(), x.data_ptr(), rows, logits.data_ptr(), ws.data_ptr(), ws.numel(),
                                               torch.cuda.current_stream(dev).cuda_stream)
        L.check(rc, "psvae_consistency_forward")
        return logits

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), **self.optimizer_cfg)

    # ---- the stand-alone trainer (embedding_classifier.py:64-100) ---------------------------------------------------
    def _tensors(self):
        return [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.fc3.weight, self.fc3.bias]

    def _step(self, batch, prefix: str, compute_grads: bool) -> torch.Tensor:
        x, y = batch
        dev = self.fc1.weight.device
        if dev.type != "cuda":
            raise RuntimeError(f"pseudo_speaker_vae_b200 runs on a B200 only (module is on {dev}); there is no CPU fallback")
        if x.dim() != 2 or x.shape[1] != self.input_dim:
            raise ValueError(f"x must be [batch, {self.input_dim}], got {tuple(x.shape)}")
        if x.device != dev:
            raise ValueError(f"x is on {x.device}, the model on {dev}")
        x = x.detach().to(torch.float32).contiguous()
        y = y.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        rows = x.shape[0]
        if rows == 0 or y.numel() != rows:
            raise ValueError(f"labels have {y.numel()} rows, x has {rows}")
        d, flat = self.flat_params(dev)
        need_grads = compute_grads and torch.is_grad_enabled() and any(t.requires_grad for t in self._tensors())
        gflat = torch.empty(int(d.total_numel), dtype=torch.float32, device=dev) if need_grads else None
        losses = torch.empty(L.NUM_LOSSES, dtype=torch.float32, device=dev)
        need = int(L.lib().psvae_embedding_classifier_workspace_bytes(C.byref(d), rows))
        if need < 0:
            raise ValueError(L.last_error())
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = L.lib().psvae_embedding_classifier_step(C.byref(d), flat.data_ptr(), L.ptr(gflat), x.data_ptr(), y.data_ptr(), rows, int(need_grads), None,
                                                         losses.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
        L.check(rc, "psvae_embedding_classifier_step")
        self.log(f"{prefix}_acc", losses[L.LOSS_CONS_ACC], sync_dist=True)
        self.log(f"{prefix}_loss", losses[L.LOSS_TOTAL], sync_dist=True)
        if not need_grads:
            return losses[L.LOSS_TOTAL].clone()
        offs = [d.w[0], d.b[0], d.w[1], d.b[1], d.w[2], d.b[2]]
        return _ClassifierLoss.apply(losses[L.LOSS_TOTAL], gflat, offs, *self._tensors())

    def training_step(self, batch, batch_idx: int) -> torch.Tensor:
        """CrossEntropyLoss(self(x), y); logs ``train_acc`` / ``train_loss`` (embedding_classifier.py:64-82)."""
        return self._step(batch, "train", True)

    def validation_step(self, batch, batch_idx: int) -> torch.Tensor:
        """The same without gradients; logs ``val_acc`` / ``val_loss`` (embedding_classifier.py:84-100)."""
        return self._step(batch, "val", False)


class _ClassifierLoss(torch.autograd.Function):
    """``loss.backward()`` for the stand-alone classifier step: the six gradient tensors were computed by the library call."""

    @staticmethod
    def forward(ctx, loss_value, gflat, offs, *params):
        ctx.gflat, ctx.offs = gflat, offs
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.needs = [p.requires_grad for p in params]
        return loss_value.clone()

    @staticmethod
    def backward(ctx, gout):
        g = ctx.gflat
        if g is None:
            raise RuntimeError("the gradients of this loss were already handed to autograd; call training_step again")
        ctx.gflat = None
        g = g * gout
        outs = []
        for off, shp, need in zip(ctx.offs, ctx.shapes, ctx.needs):
            n = 1
            for v in shp:
                n *= v
            outs.append(g[off:off + n].view(shp) if need else None)
        return (None, None, None) + tuple(outs)
