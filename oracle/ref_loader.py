"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (this container only).

Imports ``/root/reference/ps_vae/{model,latent_classifier,lightning,inference}.py`` as they lie on
disk, after putting ~30 lines of in-memory stand-ins for ``pytorch_lightning`` and ``torchmetrics``
into ``sys.modules`` (neither is installed here and there is no network).  Nothing from
the reference is copied into this repo.  ``/root/reference`` does not exist on the GPU box, so this
module is used only by ``oracle/make_golden.py`` (fixture generation) and by ``-m "not gpu"`` tests
that skip when the reference is absent.  It is never imported by the product package.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PSVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ps_vae", "model.py"))


class _HParams(dict):
    """`hparams.model['latent_dim']` and `hparams["optimizer"]` both work (inference.py:22, lightning.py:205)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _install_stubs() -> None:
    if "pytorch_lightning" in sys.modules and "torchmetrics" in sys.modules:
        return

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self._hparams = _HParams()
            self.logged = {}

        def save_hyperparameters(self, *args, **kwargs):
            import inspect

            frame = inspect.currentframe().f_back
            local = frame.f_locals
            hp = {}
            for name, val in local.items():
                if name in ("self", "__class__"):
                    continue
                if name == "hparams" and isinstance(val, dict):
                    hp.update(val)
                else:
                    hp[name] = val
            self._hparams = _HParams(hp)

        @property
        def hparams(self):
            return self._hparams

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:  # pragma: no cover
                return torch.device("cpu")

        def log(self, name, value, **kw):
            self.logged[name] = value.detach().clone() if torch.is_tensor(value) else value

    class Callback:  # noqa: D401
        pass

    def seed_everything(seed, workers=False):
        torch.manual_seed(seed)

    pl.LightningModule = LightningModule
    pl.Callback = Callback
    pl.seed_everything = seed_everything
    sys.modules["pytorch_lightning"] = pl

    tm = types.ModuleType("torchmetrics")

    class Accuracy(nn.Module):
        def __init__(self, task="multiclass", num_classes=None, **kw):
            super().__init__()
            self.num_classes = num_classes

        def forward(self, preds, target):
            if preds.ndim == target.ndim + 1:
                preds = preds.argmax(dim=-1)
            return (preds == target).float().mean()

    tm.Accuracy = Accuracy
    sys.modules["torchmetrics"] = tm

    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except Exception:  # pragma: no cover
            tq = types.ModuleType("tqdm")
            tq.tqdm = lambda it, **k: it
            sys.modules["tqdm"] = tq


def load_reference():
    """Return a namespace with the reference's live classes/functions (unmodified code)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    model = importlib.import_module("ps_vae.model")
    latent_classifier = importlib.import_module("ps_vae.latent_classifier")
    lightning = importlib.import_module("ps_vae.lightning")
    inference = importlib.import_module("ps_vae.inference")
    ns = types.SimpleNamespace(
        VAEModel=model.VAEModel,
        LatentClassifier=latent_classifier.LatentClassifier,
        PseudoSpeakerVAE=lightning.PseudoSpeakerVAE,
        unconditional_synthesis=inference.unconditional_synthesis,
        conditional_synthesis=inference.conditional_synthesis,
        inference_module=inference,
        model_module=model,
    )
    return ns


@contextlib.contextmanager
def injected_normals(tensors):
    """Make ``torch.randn`` / ``torch.randn_like`` return the given tensors in order (F7).

    Forward: one ``randn_like(sigma)`` draw (model.py:57).  conditional_synthesis: ``randn((N,L))``
    then one ``randn_like(z)`` per step (inference.py:73,95)."""
    queue = list(tensors)
    orig_randn, orig_randn_like = torch.randn, torch.randn_like

    def _pop(shape, dtype, requires_grad=False):
        t = queue.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        t = t.clone().to(dtype)
        if requires_grad:
            t.requires_grad_(True)
        return t

    def randn(*size, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        return _pop(size, kw.get("dtype") or torch.get_default_dtype(), kw.get("requires_grad", False))

    def randn_like(t, **kw):
        return _pop(t.shape, t.dtype)

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_randn_like
    assert not queue, f"{len(queue)} injected normals were not consumed"
