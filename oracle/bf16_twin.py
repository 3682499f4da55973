"""TEST INFRASTRUCTURE ONLY -- the bf16-emulating twin of ``oracle/ps_vae_oracle.train_loss_and_grads``.

The tensor-core mode of the library (PSVAE_BF16) rounds to bf16 at fixed places: the operand copies of x and of every
weight matrix, every stored hidden activation, z, sigma*eps/2, and every activation gradient that becomes the A operand of
the next dgrad / wgrad.  Against the fp64 twin of the reference those roundings show up as 1e-3 .. 1e-1 on the gradients
(a ReLU unit whose pre-activation lies within bf16 rounding of zero takes the other subgradient), which is too wide a bar to
catch an epilogue bug.  This file restates the SAME arithmetic as ps_vae_oracle (ps_vae/model.py:38-63,
ps_vae/lightning.py:67-131 and its autograd) with a round-to-nearest-even bf16 rounding at exactly those places and
fp64 accumulation everywhere else, so the bf16 kernels can be held to ~1e-3 (what is left: fp32 vs fp64 accumulation order
moving an element across a bf16 rounding boundary, 2^-9 on that element).

Where the kernels round (pseudo_speaker_vae_b200/csrc):
  forward   x, W -> bf16 (cast_bf16_kernel, the Adam pass's shadow copy); h = bf16(relu(acc + b)) (EpiBiasAct, the ReLU mask is
            taken from the fp32 pre-activation); mu, log_sigma, x_hat, logits stay fp32; z, sigma*eps/2 -> bf16
  loss      every term in fp32 from the fp32 mu / log_sigma / x_hat and the caller's x
  backward  d x_hat -> bf16 (EpiMse / recon_rows_kernel); each hidden-layer gradient g = bf16((dY W) * relu') (EpiActGrad); bias
            gradients of hidden layers and of the last decoder layer are column sums of the ROUNDED gradient; dz fp32;
            d mu, d log_sigma computed in fp32 (d log_sigma from the bf16 stash of sigma*eps/2), their column sums (= the head
            biases' gradients) taken BEFORE rounding, then rounded to bf16 as the operands of the head wgrad / dgrad
  classifier / consistency classifier: fp32 throughout (CUDA cores)

Only ``tests/`` and ``__graft_entry__.smoke()`` import this module.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import ps_vae_oracle as O


def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float64 (values exactly representable in bf16)."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32)
    out = np.where(np.isfinite(f), out, f)
    return out.astype(np.float64)


def _f32(a: np.ndarray) -> np.ndarray:
    """Round to fp32 (what leaves TMEM), kept as float64."""
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def _mlp_fwd(params, prefix, a_in):
    """Linear/ReLU chain with bf16 operands and bf16 stored activations.  Returns (fp32 output of the last Linear, cache) where
    cache[j] = (bf16 input of Linear j, ReLU mask of its output or None)."""
    idx = O._linear_indices(params, prefix)
    a = a_in
    cache = []
    for j, i in enumerate(idx):
        Wb = bf16_round(params[f"{prefix}.{i}.weight"])
        b = np.asarray(params[f"{prefix}.{i}.bias"], dtype=np.float64)
        u = _f32(_f32(a @ Wb.T) + b)
        if j + 1 < len(idx):
            cache.append((a, u > 0))
            a = bf16_round(np.maximum(u, 0))
        else:
            cache.append((a, None))
            a = u
    return a, cache


def _mlp_bwd(params, prefix, cache, dout_b, grads, first_bias_from: Optional[np.ndarray] = None, need_dx: bool = False):
    """dout_b: bf16-valued d loss / d (output of the last Linear).  first_bias_from: the fp32 (un-rounded) gradient whose column
    sums are the last Linear's bias gradient (the latent backward kernel sums before rounding); None: sum the rounded values."""
    idx = O._linear_indices(params, prefix)
    d = dout_b
    dx = None
    for j in reversed(range(len(idx))):
        i = idx[j]
        a_in, _ = cache[j]
        grads[f"{prefix}.{i}.weight"] = d.T @ a_in
        src = first_bias_from if (first_bias_from is not None and j + 1 == len(idx)) else d
        grads[f"{prefix}.{i}.bias"] = src.sum(axis=0)
        if j > 0:
            Wb = bf16_round(params[f"{prefix}.{i}.weight"])
            mask = cache[j - 1][1]
            d = bf16_round(_f32(d @ Wb) * mask)
        elif need_dx:
            Wb = bf16_round(params[f"{prefix}.{i}.weight"])
            dx = _f32(d @ Wb)
    return dx


def train_loss_and_grads_bf16(params, x, y, eps, *, kl_loss_weight=1.0, classifier_loss_weight=1.0, normalize_decoder=False, use_cos_loss=False,
                              classifier_activation="relu", compute_grads=True, consistency_params=None, consistency_loss_weight=1.0):
    """Same contract as ``ps_vae_oracle.train_loss_and_grads`` (float64 in / out), bf16 roundings where the PSVAE_BF16 kernels round.
    ``x`` is the caller's input as the kernels see it for the loss (fp32 values, or bf16 values when the batch comes from a bf16 store)."""
    params = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    x = np.asarray(x, dtype=np.float64)
    eps = np.asarray(eps, dtype=np.float64)
    B, D = x.shape
    xa = bf16_round(x)
    mu, c_mu = _mlp_fwd(params, "model.encoder_mu", xa)
    ls, c_ls = _mlp_fwd(params, "model.encoder_sigma", xa)
    sigma = np.exp(0.5 * ls)
    z_b = bf16_round(mu + sigma * eps)
    hs_b = bf16_round(0.5 * sigma * eps)
    u, c_dec = _mlp_fwd(params, "model.decoder", z_b)
    if normalize_decoder:
        x_hat, den = O._normalize_rows(u)
    else:
        x_hat, den = u, None
    has_clf = any(k.startswith("classifier.") for k in params)
    scal = {}
    grads = {} if compute_grads else None

    if use_cos_loss:
        EPS = 1e-12
        dot = (x_hat * x).sum(axis=1)
        m1 = (x_hat * x_hat).sum(axis=1) + EPS
        m2 = (x * x).sum(axis=1) + EPS
        dn = np.sqrt(m1 * m2)
        cos = dot / dn
        recon = (1 - cos).mean()
        dxh = -(x / dn[:, None] - (cos / m1)[:, None] * x_hat) / B
    else:
        diff = x_hat - x
        recon = (diff * diff).mean() / 10.0
        dxh = diff * (2.0 / (B * D * 10.0))
    els = np.exp(ls)
    kl = -0.5 * (1 + ls - mu * mu - els).sum(axis=-1).mean()

    clf_loss = 0.0
    dmu_clf = 0
    if has_clf:
        logits, ccache = O.classifier_forward(params, mu, classifier_activation)
        if isinstance(logits, dict):
            dlog = {}
            n = len(logits)
            for name, lg in logits.items():
                l_, d_ = O.cross_entropy(lg, y[name])
                clf_loss = clf_loss + l_
                dlog[name] = d_ * (classifier_loss_weight / n)
                scal[f"classifier_acc_{name}"] = (lg.argmax(-1) == y[name]).mean()
            clf_loss = clf_loss / n
        else:
            clf_loss, dlog = O.cross_entropy(logits, y)
            dlog = dlog * classifier_loss_weight
            scal["classifier_acc"] = (logits.argmax(-1) == y).mean()
        if compute_grads:
            dmu_clf = O.classifier_backward(params, ccache, dlog, grads, classifier_activation)

    cons_loss = 0.0
    if consistency_params is not None:
        cp = {k: np.asarray(v, dtype=np.float64) for k, v in consistency_params.items()}
        logits_c, c_cons = O.embedding_classifier_forward(cp, x_hat)
        cons_loss, dlog_c = O.cross_entropy(logits_c, y)
        scal["consistency_acc"] = (logits_c.argmax(-1) == y).mean()
        if compute_grads:
            dxh = dxh + O.embedding_classifier_input_grad(cp, c_cons, dlog_c * consistency_loss_weight)

    total = recon + kl_loss_weight * kl + classifier_loss_weight * clf_loss + consistency_loss_weight * cons_loss
    scal.update(loss=total, recon_loss=recon, kl_loss=kl, classifier_loss=clf_loss, consistency_loss=cons_loss)
    outputs = dict(x_hat=x_hat, mu=mu, log_sigma=ls, z=z_b)
    if not compute_grads:
        return scal, outputs, None

    if normalize_decoder:
        du = (dxh - x_hat * (x_hat * dxh).sum(axis=1, keepdims=True)) / den
    else:
        du = dxh
    du_b = bf16_round(du)
    dz = _mlp_bwd(params, "model.decoder", c_dec, du_b, grads, need_dx=True)
    dmu = _f32(dz + (kl_loss_weight / B) * mu + dmu_clf)
    dls = _f32(dz * hs_b + (kl_loss_weight * 0.5 / B) * np.expm1(ls))
    _mlp_bwd(params, "model.encoder_mu", c_mu, bf16_round(dmu), grads, first_bias_from=dmu)
    _mlp_bwd(params, "model.encoder_sigma", c_ls, bf16_round(dls), grads, first_bias_from=dls)
    outputs.update(dz=dz, dmu=dmu, dls=dls)
    return scal, outputs, grads
