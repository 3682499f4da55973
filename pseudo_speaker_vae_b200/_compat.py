"""The slice of ``pytorch_lightning`` / ``torchmetrics`` the reference touches (ps_vae/lightning.py:1,8,33,38), for images where
neither package is installed: the real ``LightningModule`` is used as the base class when it is importable."""
from __future__ import annotations

import inspect
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as _pl

    _Base = _pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _pl = None
    HAVE_LIGHTNING = False

    class _AttrDict(dict):
        """``hparams.model['latent_dim']`` and ``hparams["optimizer"]`` both work (inference.py:22, lightning.py:205)."""

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v

    class _Base(nn.Module):
        """The slice of LightningModule the reference touches: save_hyperparameters, hparams, log, device,
        load_from_checkpoint (checkpoint dict keys ``hyper_parameters`` / ``state_dict``)."""

        def __init__(self, *a, **k):
            super().__init__()
            self._hparams = _AttrDict()
            self.logged: Dict[str, Any] = {}

        def save_hyperparameters(self, *args, **kwargs):
            frame = inspect.currentframe().f_back
            hp = {}
            for name, val in frame.f_locals.items():
                if name in ("self", "__class__"):
                    continue
                if isinstance(val, dict) and name in ("hparams", "kwargs"):
                    hp.update(val)
                else:
                    hp[name] = val
            self._hparams = _AttrDict(hp)

        @property
        def hparams(self):
            return self._hparams

        @property
        def device(self) -> torch.device:
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, name, value, **kw):
            self.logged[name] = value

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **overrides):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(overrides)
            module = cls(**hp)
            module.load_state_dict(ckpt["state_dict"])
            return module


class _Accuracy(nn.Module):
    """Stand-in for ``torchmetrics.Accuracy(task='multiclass')``: mean(argmax == y).  (The fused step computes the
    same number in-kernel; this object exists so ``module.accuracy`` keeps its place in the attribute surface.)"""

    def __init__(self, task: str = "multiclass", num_classes: Optional[int] = None):
        super().__init__()
        self.task, self.num_classes = task, num_classes

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return (preds.argmax(dim=-1) == target).float().mean()
