"""Shared helpers of the GPU parity tests: build models from the golden-case configs, talk to the C-ABI."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pseudo_speaker_vae_b200 import _lib as L  # noqa: E402
import pseudo_speaker_vae_b200 as P  # noqa: E402
from tests.golden_util import case_consistency_params, case_params  # noqa: E402

DEV = "cuda:0"


def stream():
    return torch.cuda.current_stream().cuda_stream


def module_from_cfg(cfg, precision="fp32", params=None, device=DEV, **extra):
    """PseudoSpeakerVAE carrying the synthetic parameters of a golden case (tests/golden_util.case_params)."""
    hp = dict(model=dict(input_dim=cfg["D"], latent_dim=cfg["L"], normalize_decoder=cfg.get("normalize_decoder", False)),
              optimizer=dict(cfg.get("optimizer", dict(lr=1e-3))), scheduler=dict(T_max=200), precision=precision,
              kl_loss_weight=cfg.get("kl_w", 1.0), classifier_loss_weight=cfg.get("clf_w", 1.0), use_cos_loss=cfg.get("use_cos_loss", False),
              consistency_loss_weight=cfg.get("cons_w", 1.0))
    if "H" in cfg:
        hp["model"].update(hidden_dim=cfg["H"], num_hidden_layers=cfg["nh"])
    if cfg.get("clf"):
        hp["classifier"] = dict(cfg["clf"])
    hp.update(extra)
    m = P.PseudoSpeakerVAE(**hp)
    if params is None:
        params = case_params(cfg, np.float32)
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("consistency_classifier.") for k in missing), (missing, unexpected)
    if cfg.get("cons"):
        # what `consistency_classifier_ckpt` does (lightning.py:44-52), with the golden case's synthetic weights instead of a checkpoint file
        c = cfg["cons"]
        ec = P.EmbeddingClassifier(input_dim=cfg["D"], num_classes=c["num_classes"], hidden_dim=c.get("hidden_dim", 128))
        ec.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in case_consistency_params(cfg, np.float32).items()})
        for p in ec.parameters():
            p.requires_grad = False
        ec.eval()
        m.consistency_classifier = ec
    return m.to(device)


def labels_to_torch(y, device=DEV):
    if isinstance(y, dict):
        return {k: torch.from_numpy(v).to(device) for k, v in y.items()}
    return torch.from_numpy(y).to(device)


def grads_dict(module):
    return {k: p.grad.detach().cpu().numpy() for k, p in module.named_parameters() if p.grad is not None}


def flat_to_dict(module, flat):
    """Slice a flat arena-layout buffer back into {state_dict key: array}."""
    hot = module.hot_path
    names = {id(p): k for k, p in module.named_parameters()}
    out = {}
    f = flat.detach().cpu().numpy()
    for p, off in hot.arena.entries:
        out[names[id(p)]] = f[off:off + p.numel()].reshape(tuple(p.shape))
    return out


def gemm_bf16(a, b, bias=None, a_mn=False, b_mn=False, relu=False, split_k=1):
    """C[M,N] = A[M,K] B[N,K]^T through psvae_gemm_bf16.  a: [M,K] float tensor, b: [N,K]; stored MN-major when asked."""
    M, K = a.shape
    N = b.shape[0]
    a16 = (a.t().contiguous() if a_mn else a.contiguous()).to(torch.bfloat16)
    b16 = (b.t().contiguous() if b_mn else b.contiguous()).to(torch.bfloat16)
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device=a.device)
    ws = torch.empty(max(1, split_k) * M * N * 4 + 256, dtype=torch.uint8, device=a.device) if split_k > 1 else None
    rc = L.lib().psvae_gemm_bf16(a16.data_ptr(), b16.data_ptr(), L.ptr(bias), c.data_ptr(), M, N, K, int(a_mn), int(b_mn), int(relu), split_k,
                                 L.ptr(ws), ws.numel() if ws is not None else 0, stream())
    L.check(rc, "psvae_gemm_bf16")
    ref = a16.t().double() if a_mn else a16.double()
    refb = b16.t().double() if b_mn else b16.double()
    r = ref @ refb.t()
    if bias is not None:
        r = r + bias.double()
    if relu:
        r = r.clamp_min(0)
    return c, r
