"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the pseudo_speaker_VAE hot path.

This file is the parity oracle for the B200 kernels.  It is a from-scratch restatement of the
reference's arithmetic, with closed-form gradients instead of autograd, and it is *pinned*: the
fixtures under ``tests/golden/`` were produced by running the unmodified reference
(``/root/reference``) in this container (``oracle/make_golden.py``) and ``tests/test_oracle_golden.py``
checks every function below against them.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package never does.

Reference lines restated (paths relative to /root/reference):
  * ``vae_forward`` / ``decode``           -- ps_vae/model.py:54-63, :65-69
  * ``classifier_forward``                 -- ps_vae/latent_classifier.py:30-70
  * ``train_loss_and_grads``               -- ps_vae/lightning.py:67-131 (loss) + autograd of it
  * ``adam_step``                          -- torch/optim/adam.py:416-547 (single-tensor path) as
                                               driven by ps_vae/lightning.py:204-205
  * ``cosine_annealing_lr``                -- torch/optim/lr_scheduler.py (CosineAnnealingLR.get_lr)
                                               as driven by ps_vae/lightning.py:206-213
  * ``unconditional_synthesis``            -- ps_vae/inference.py:10-27
  * ``langevin`` / ``conditional_synthesis``-- ps_vae/inference.py:29-110
  * label tables / target parsing          -- ps_vae/utils.py:82-136, ps_vae/inference.py:128-132

Parameters travel as a ``dict[str, np.ndarray]`` keyed by the reference's own state_dict names
(``model.encoder_mu.0.weight`` ... ``classifier.layers.0.weight`` / ``classifier.output_layers.<label>.weight``).
``Linear`` is ``a @ W.T + b`` with ``W:[out,in]``.  The hidden width / depth are inferred from the
keys, so the widened twin of BASELINE config 5 goes through the same code.
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

Params = Dict[str, np.ndarray]

# --------------------------------------------------------------------------------------------
# label tables (ps_vae/utils.py:82-136) and CLI target parsing (ps_vae/inference.py:128-132)
# --------------------------------------------------------------------------------------------
CV_AGE_TO_LABEL = {
    "teens": 0, "twenties": 0,
    "thirties": 1, "fourties": 1, "fifties": 1,
    "sixties": 2, "seventies": 2, "eighties": 2, "nineties": 2,
}
CV_GENDER_TO_LABEL = {"male": 0, "female": 1, "other": 2}
VCTK_GENDER_TO_LABEL = {"M": 0, "F": 1}


def map_cv_age_to_label(age) -> int:
    return CV_AGE_TO_LABEL.get(age, -1)


def map_cv_gender_to_label(gender) -> int:
    return CV_GENDER_TO_LABEL.get(gender, -1)


def map_vctk_gender_to_label(gender) -> int:
    return VCTK_GENDER_TO_LABEL.get(gender, -1)


def parse_classifier_target(text: str) -> Union[int, dict]:
    """JSON first, then int (inference.py:128-132)."""
    try:
        return json.loads(text)
    except json.JSONDecodeError:
        return int(text)


def sample_filename(i: int) -> str:
    """Row i of the sampled batch <-> file name (inference.py:154-156)."""
    return f"sample_{i}.pt"


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def _linear_indices(params: Params, prefix: str) -> List[int]:
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in params if k.startswith(prefix + ".") and k.endswith(".weight")})
    return idx


def _act(name: str, u: np.ndarray) -> np.ndarray:
    if name == "relu":
        return np.maximum(u, 0)
    if name == "tanh":
        return np.tanh(u)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-u))
    if name == "leaky_relu":
        return np.where(u > 0, u, u * u.dtype.type(0.01))
    raise ValueError(f"Unsupported activation: {name}")


def _act_grad(name: str, u: np.ndarray, a: np.ndarray) -> np.ndarray:
    """d act(u) / du given pre-activation u and post-activation a."""
    if name == "relu":
        return (u > 0).astype(u.dtype)
    if name == "tanh":
        return 1 - a * a
    if name == "sigmoid":
        return a * (1 - a)
    if name == "leaky_relu":
        return np.where(u > 0, u.dtype.type(1), u.dtype.type(0.01))
    raise ValueError(name)


def mlp_forward(params: Params, prefix: str, x: np.ndarray, act: str = "relu"):
    """nn.Sequential(Linear, act, Linear, act, ..., Linear) -- model.py:14-36.

    Returns (output, cache) where cache = list of (input_to_linear, pre_activation)."""
    idx = _linear_indices(params, prefix)
    a = x
    cache = []
    for j, i in enumerate(idx):
        W = params[f"{prefix}.{i}.weight"]
        b = params[f"{prefix}.{i}.bias"]
        u = a @ W.T + b
        cache.append((a, u))
        a = _act(act, u) if j + 1 < len(idx) else u
    return a, cache


def mlp_backward(params: Params, prefix: str, cache, dout: np.ndarray, grads: Optional[Params], act: str = "relu", need_dx: bool = True):
    """Backward of mlp_forward.  Accumulates weight/bias grads into ``grads`` (if given); returns d input."""
    idx = _linear_indices(params, prefix)
    d = dout
    for j in reversed(range(len(idx))):
        i = idx[j]
        a_in, u = cache[j]
        if j + 1 < len(idx):
            d = d * _act_grad(act, u, _act(act, u))
        if grads is not None:
            grads[f"{prefix}.{i}.weight"] = grads.get(f"{prefix}.{i}.weight", 0) + d.T @ a_in
            grads[f"{prefix}.{i}.bias"] = grads.get(f"{prefix}.{i}.bias", 0) + d.sum(axis=0)
        if j > 0 or need_dx:
            d = d @ params[f"{prefix}.{i}.weight"]
    return d


def _normalize_rows(u: np.ndarray):
    """F.normalize(u, p=2, dim=1): u / max(||u||_2, 1e-12) -- model.py:60-61."""
    n = np.sqrt((u * u).sum(axis=1, keepdims=True))
    den = np.maximum(n, u.dtype.type(1e-12))
    return u / den, den


# --------------------------------------------------------------------------------------------
# VAE forward / decode (ps_vae/model.py)
# --------------------------------------------------------------------------------------------
def vae_forward(params: Params, x: np.ndarray, eps: np.ndarray, normalize_decoder: bool = False):
    """x -> (x_hat, mu, log_sigma); eps is the injected randn_like(sigma) draw (model.py:54-63)."""
    mu, c_mu = mlp_forward(params, "model.encoder_mu", x)
    ls, c_ls = mlp_forward(params, "model.encoder_sigma", x)
    sigma = np.exp(x.dtype.type(0.5) * ls)
    z = mu + sigma * eps
    u, c_dec = mlp_forward(params, "model.decoder", z)
    if normalize_decoder:
        x_hat, den = _normalize_rows(u)
    else:
        x_hat, den = u, None
    cache = dict(c_mu=c_mu, c_ls=c_ls, c_dec=c_dec, sigma=sigma, z=z, u=u, den=den)
    return x_hat, mu, ls, cache


def decode(params: Params, z: np.ndarray, normalize_decoder: bool = False) -> np.ndarray:
    """model.py:65-69."""
    u, _ = mlp_forward(params, "model.decoder", z)
    if normalize_decoder:
        u, _ = _normalize_rows(u)
    return u


# --------------------------------------------------------------------------------------------
# latent classifier (ps_vae/latent_classifier.py)
# --------------------------------------------------------------------------------------------
def classifier_heads(params: Params) -> List[str]:
    """Names of the multi-label heads, in ModuleDict insertion order as stored in the dict."""
    heads = []
    for k in params:
        if k.startswith("classifier.output_layers.") and k.endswith(".weight"):
            heads.append(k[len("classifier.output_layers."):-len(".weight")])
    return heads


def classifier_forward(params: Params, h: np.ndarray, activation: str = "relu"):
    """Trunk ``classifier.layers.*`` (Linear / act alternating) then optional dict of heads
    (latent_classifier.py:58-70).  In single-label mode the last trunk Linear is the output.

    Returns (logits | {label: logits}, cache)."""
    heads = classifier_heads(params)
    idx = _linear_indices(params, "classifier.layers")
    a = h
    cache = []
    for j, i in enumerate(idx):
        W = params[f"classifier.layers.{i}.weight"]
        b = params[f"classifier.layers.{i}.bias"]
        u = a @ W.T + b
        last_single = (not heads) and j + 1 == len(idx)
        cache.append((a, u, last_single))
        a = u if last_single else _act(activation, u)
    if not heads:
        return a, dict(trunk=cache, feat=None)
    out = {}
    for name in heads:
        out[name] = a @ params[f"classifier.output_layers.{name}.weight"].T + params[f"classifier.output_layers.{name}.bias"]
    return out, dict(trunk=cache, feat=a)


def classifier_backward(params: Params, cache, dlogits, grads: Optional[Params], activation: str = "relu") -> np.ndarray:
    """Given d loss / d logits (array or dict per head) return d loss / d input; accumulate param grads."""
    heads = classifier_heads(params)
    if heads:
        feat = cache["feat"]
        d = 0
        for name in heads:
            dl = dlogits.get(name)
            if dl is None:
                continue
            W = params[f"classifier.output_layers.{name}.weight"]
            if grads is not None:
                grads[f"classifier.output_layers.{name}.weight"] = dl.T @ feat
                grads[f"classifier.output_layers.{name}.bias"] = dl.sum(axis=0)
            d = d + dl @ W
    else:
        d = dlogits
    idx = _linear_indices(params, "classifier.layers")
    for j in reversed(range(len(idx))):
        i = idx[j]
        a_in, u, last_single = cache["trunk"][j]
        if not last_single:
            d = d * _act_grad(activation, u, _act(activation, u))
        if grads is not None:
            grads[f"classifier.layers.{i}.weight"] = d.T @ a_in
            grads[f"classifier.layers.{i}.bias"] = d.sum(axis=0)
        d = d @ params[f"classifier.layers.{i}.weight"]
    return d


def embedding_classifier_forward(params: Params, x: np.ndarray):
    """EmbeddingClassifier.forward (ps_vae/embedding_classifier/embedding_classifier.py:50-62):
    fc3(relu(fc2(relu(fc1(x))))).  Returns (logits, cache)."""
    a1 = np.maximum(x @ params["fc1.weight"].T + params["fc1.bias"], 0)
    a2 = np.maximum(a1 @ params["fc2.weight"].T + params["fc2.bias"], 0)
    logits = a2 @ params["fc3.weight"].T + params["fc3.bias"]
    return logits, (a1, a2)


def embedding_classifier_input_grad(params: Params, cache, dlogits: np.ndarray) -> np.ndarray:
    """d loss / d x of the frozen EmbeddingClassifier (no parameter gradients: lightning.py:48-49 freezes it)."""
    a1, a2 = cache
    d2 = (dlogits @ params["fc3.weight"]) * (a2 > 0)
    d1 = (d2 @ params["fc2.weight"]) * (a1 > 0)
    return d1 @ params["fc1.weight"]


def embedding_classifier_param_shapes(input_dim: int, num_classes: int, hidden_dim: int = 128):
    """state_dict keys / shapes of EmbeddingClassifier (embedding_classifier.py:29-32)."""
    return [("fc1.weight", (hidden_dim, input_dim)), ("fc1.bias", (hidden_dim,)), ("fc2.weight", (hidden_dim, hidden_dim)),
            ("fc2.bias", (hidden_dim,)), ("fc3.weight", (num_classes, hidden_dim)), ("fc3.bias", (num_classes,))]


def log_softmax(logits: np.ndarray) -> np.ndarray:
    m = logits.max(axis=-1, keepdims=True)
    s = logits - m
    return s - np.log(np.exp(s).sum(axis=-1, keepdims=True))


def cross_entropy(logits: np.ndarray, y: np.ndarray):
    """F.cross_entropy(logits, y) with mean reduction; returns (loss, dloss/dlogits)."""
    B = logits.shape[0]
    lp = log_softmax(logits)
    loss = -lp[np.arange(B), y].mean(dtype=logits.dtype)
    p = np.exp(lp)
    p[np.arange(B), y] -= 1
    return loss, p / logits.dtype.type(B)


# --------------------------------------------------------------------------------------------
# training step: loss (lightning.py:67-131) + closed-form backward
# --------------------------------------------------------------------------------------------
def train_loss_and_grads(
    params: Params,
    x: np.ndarray,
    y,
    eps: np.ndarray,
    *,
    kl_loss_weight: float = 1.0,
    classifier_loss_weight: float = 1.0,
    normalize_decoder: bool = False,
    use_cos_loss: bool = False,
    classifier_activation: str = "relu",
    compute_grads: bool = True,
    consistency_params: Optional[Params] = None,
    consistency_loss_weight: float = 1.0,
):
    """Returns (scalars: dict, outputs: dict, grads: dict|None).

    consistency_params: the frozen EmbeddingClassifier (``fc1/fc2/fc3`` state-dict keys) of lightning.py:44-52; its
    cross entropy on x_hat is added with ``consistency_loss_weight`` (lightning.py:100-108,119-124) and its
    gradient flows into the decoder through x_hat (the classifier itself gets no gradient).

    scalars: loss, recon_loss, kl_loss, classifier_loss, classifier_acc (names follow the ``train_*`` /
    ``val_*`` metric names of lightning.py:82-83,127-129).  y: int array [B] (single label) or
    {label: int array} for a multi-head classifier (there the CE is summed over heads and divided by the
    number of heads -- the reference's own multi-label Lightning branch is broken)."""
    dt = x.dtype.type
    B, D = x.shape
    x_hat, mu, ls, c = vae_forward(params, x, eps, normalize_decoder)
    has_clf = any(k.startswith("classifier.") for k in params)
    scal = {}
    grads: Optional[Params] = {} if compute_grads else None

    # reconstruction term (lightning.py:110-113)
    if use_cos_loss:
        EPS = dt(1e-12)
        dot = (x_hat * x).sum(axis=1)
        m1 = (x_hat * x_hat).sum(axis=1) + EPS
        m2 = (x * x).sum(axis=1) + EPS
        den = np.sqrt(m1 * m2)
        cos = dot / den
        recon = (1 - cos).mean(dtype=x.dtype)
        dxh = -(x / den[:, None] - (cos / m1)[:, None] * x_hat) / dt(B)
    else:
        diff = x_hat - x
        recon = (diff * diff).mean(dtype=x.dtype) / dt(10)
        dxh = diff * dt(2.0 / (B * D * 10.0))
    # KL (lightning.py:115-117)
    els = np.exp(ls)
    kl = dt(-0.5) * (1 + ls - mu * mu - els).sum(axis=-1).mean(dtype=x.dtype)

    clf_loss = dt(0)
    dmu_clf = 0
    if has_clf:
        logits, ccache = classifier_forward(params, mu, classifier_activation)
        if isinstance(logits, dict):
            dlog = {}
            n = len(logits)
            for name, lg in logits.items():
                l_, d_ = cross_entropy(lg, y[name])
                clf_loss = clf_loss + l_
                dlog[name] = d_ * dt(classifier_loss_weight / n)
                scal[f"classifier_acc_{name}"] = (lg.argmax(-1) == y[name]).mean()
            clf_loss = clf_loss / dt(n)
        else:
            clf_loss, dlog = cross_entropy(logits, y)
            dlog = dlog * dt(classifier_loss_weight)
            scal["classifier_acc"] = (logits.argmax(-1) == y).mean()
        if compute_grads:
            dmu_clf = classifier_backward(params, ccache, dlog, grads, classifier_activation)

    cons_loss = dt(0)
    if consistency_params is not None:
        # lightning.py:100-108: CE(consistency_classifier(x_hat), y); y is the single-label tensor
        logits_c, c_cons = embedding_classifier_forward(consistency_params, x_hat)
        cons_loss, dlog_c = cross_entropy(logits_c, y)
        scal["consistency_acc"] = (logits_c.argmax(-1) == y).mean()
        if compute_grads:
            dxh = dxh + embedding_classifier_input_grad(consistency_params, c_cons, dlog_c * dt(consistency_loss_weight))

    total = recon + dt(kl_loss_weight) * kl + dt(classifier_loss_weight) * clf_loss + dt(consistency_loss_weight) * cons_loss
    scal.update(loss=total, recon_loss=recon, kl_loss=kl, classifier_loss=clf_loss, consistency_loss=cons_loss)
    outputs = dict(x_hat=x_hat, mu=mu, log_sigma=ls, z=c["z"])
    if not compute_grads:
        return scal, outputs, None

    # backward
    if normalize_decoder:
        du = (dxh - x_hat * (x_hat * dxh).sum(axis=1, keepdims=True)) / c["den"]
    else:
        du = dxh
    dz = mlp_backward(params, "model.decoder", c["c_dec"], du, grads)
    dmu = dz + dt(kl_loss_weight / B) * mu + dmu_clf
    dls = dz * eps * (dt(0.5) * c["sigma"]) + dt(kl_loss_weight * 0.5 / B) * (els - 1)
    mlp_backward(params, "model.encoder_mu", c["c_mu"], dmu, grads, need_dx=False)
    mlp_backward(params, "model.encoder_sigma", c["c_ls"], dls, grads, need_dx=False)
    outputs.update(dz=dz, dmu=dmu, dls=dls)
    return scal, outputs, grads


def vae_backward(params: Params, x: np.ndarray, eps: np.ndarray, g_x_hat=None, g_mu=None, g_log_sigma=None, normalize_decoder: bool = False) -> Params:
    """Gradients of an arbitrary scalar loss w.r.t. the VAE's parameters, given d loss / d (x_hat, mu, log_sigma) of
    ``VAEModel.forward`` (model.py:38-63) -- what autograd does when a caller builds a loss of its own on the outputs.
    Same chain as the backward half of ``train_loss_and_grads`` without any built-in loss term."""
    x_hat, mu, ls, c = vae_forward(params, x, eps, normalize_decoder)
    zero = lambda like: np.zeros_like(like)
    dxh = zero(x_hat) if g_x_hat is None else g_x_hat
    if normalize_decoder:
        du = (dxh - x_hat * (x_hat * dxh).sum(axis=1, keepdims=True)) / c["den"]
    else:
        du = dxh
    grads: Params = {}
    dz = mlp_backward(params, "model.decoder", c["c_dec"], du, grads)
    dmu = dz + (0 if g_mu is None else g_mu)
    dls = dz * eps * (x.dtype.type(0.5) * c["sigma"]) + (0 if g_log_sigma is None else g_log_sigma)
    mlp_backward(params, "model.encoder_mu", c["c_mu"], dmu, grads, need_dx=False)
    mlp_backward(params, "model.encoder_sigma", c["c_ls"], dls, grads, need_dx=False)
    return grads


# --------------------------------------------------------------------------------------------
# Adam + cosine LR (torch.optim semantics, driven from lightning.py:204-214)
# --------------------------------------------------------------------------------------------
def adam_scalars(step: int, lr: float, beta1: float, beta2: float):
    """Host-side scalars of torch's single-tensor Adam (torch/optim/adam.py:476-547): computed in
    Python doubles exactly as torch does, then used as fp32 scalars."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    return lr / bc1, bc2 ** 0.5


def adam_step(p, g, m, v, step: int, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One torch.optim.Adam update of one tensor; ``step`` is the 1-based step count AFTER increment.
    Returns new (p, m, v).  amsgrad/maximize are not restated (unused by the reference)."""
    dt = p.dtype.type
    b1, b2 = betas
    if weight_decay != 0:
        g = g + dt(weight_decay) * p
    m = m + (g - m) * dt(1 - b1)                      # exp_avg.lerp_(grad, 1-beta1)
    v = v * dt(b2) + dt(1 - b2) * g * g               # mul_(beta2).addcmul_(g, g, 1-beta2)
    step_size, bc2_sqrt = adam_scalars(step, lr, b1, b2)
    denom = np.sqrt(v) / dt(bc2_sqrt) + dt(eps)
    p = p + dt(-step_size) * (m / denom)              # addcdiv_(m, denom, value=-step_size)
    return p, m, v


def cosine_annealing_lr(base_lr: float, T_max: int, eta_min: float = 0.0, epochs: int = 1) -> List[float]:
    """lr after 0..epochs scheduler.step() calls, following CosineAnnealingLR.get_lr's recursion."""
    lrs = [base_lr]
    lr = base_lr
    for last_epoch in range(1, epochs + 1):
        if (last_epoch - 1 - T_max) % (2 * T_max) == 0:
            lr = lr + (base_lr - eta_min) * (1 - math.cos(math.pi / T_max)) / 2
        else:
            lr = (1 + math.cos(math.pi * last_epoch / T_max)) / (1 + math.cos(math.pi * (last_epoch - 1) / T_max)) * (lr - eta_min) + eta_min
        lrs.append(lr)
    return lrs


# --------------------------------------------------------------------------------------------
# sampling (ps_vae/inference.py)
# --------------------------------------------------------------------------------------------
def unconditional_synthesis(params: Params, z: np.ndarray, normalize_decoder: bool = False) -> np.ndarray:
    """inference.py:22-25 with the randn draw z injected."""
    return decode(params, z, normalize_decoder)


def _select_log_prob(logits: np.ndarray, target: int):
    """_get_classifer_probs (inference.py:56-70).  The 1-logit branch is degenerate in the reference
     and is rejected here."""
    if logits.shape[1] == 1:
        raise NotImplementedError("1-logit binary classifier branch is degenerate in the reference (SURVEY F11)")
    lp = log_softmax(logits)
    return lp[:, target], lp


def langevin_grad(params: Params, z: np.ndarray, target, activation: str = "relu"):
    """grad_z [ log p(y|z) + log p(z) ]  and the two scalars the reference prints
    (inference.py:80-93,103): closed form  W^T(onehot - softmax) - z  through the classifier chain."""
    logits, ccache = classifier_forward(params, z, activation)
    N = z.shape[0]
    if isinstance(logits, dict):
        assert isinstance(target, dict), "classifier_target must be a dict for multi-label classifier"
        log_p_y = 0
        dlog = {}
        for label, t in target.items():
            sel, lp = _select_log_prob(logits[label], int(t))
            log_p_y = log_p_y + sel
            d = -np.exp(lp)
            d[np.arange(N), int(t)] += 1
            dlog[label] = d
    else:
        sel, lp = _select_log_prob(logits, int(target))
        log_p_y = sel
        dlog = -np.exp(lp)
        dlog[np.arange(N), int(target)] += 1
    g = classifier_backward(params, ccache, dlog, None, activation) - z
    log_p_z = z.dtype.type(-0.5) * (z * z).sum(axis=1)
    return g, log_p_y, log_p_z


def langevin(params: Params, z0: np.ndarray, noises: Sequence[np.ndarray], target, step_size: float = 0.01,
             noise_weight: float = 1.0, activation: str = "relu", return_history: bool = False):
    """The loop of inference.py:77-103 with z0 and the per-step noise injected."""
    dt = z0.dtype.type
    z = z0.copy()
    hist = []
    for noise in noises:
        g, _, _ = langevin_grad(params, z, target, activation)
        z = z + dt(0.5 * step_size ** 2) * g + dt(step_size * noise_weight) * noise
        if return_history:
            hist.append(z.copy())
    return (z, hist) if return_history else z


def conditional_synthesis(params: Params, z0, noises, target, step_size=0.01, noise_weight=1.0,
                          normalize_decoder=False, activation="relu", return_history=False):
    out = langevin(params, z0, noises, target, step_size, noise_weight, activation, return_history)
    z, hist = out if return_history else (out, None)
    x_hat = decode(params, z, normalize_decoder)
    return (x_hat, hist) if return_history else x_hat


# --------------------------------------------------------------------------------------------
# deterministic synthetic parameters (shared by fixtures, tests and bench; numpy PCG64 is stable)
# --------------------------------------------------------------------------------------------
def vae_param_shapes(input_dim=256, latent_dim=64, hidden_dim=512, num_hidden_layers=2) -> List[Tuple[str, Tuple[int, ...]]]:
    shapes = []

    def mlp(prefix, d_in, d_out):
        dims = [d_in] + [hidden_dim] * num_hidden_layers + [d_out]
        for j in range(len(dims) - 1):
            shapes.append((f"{prefix}.{2 * j}.weight", (dims[j + 1], dims[j])))
            shapes.append((f"{prefix}.{2 * j}.bias", (dims[j + 1],)))

    mlp("model.encoder_mu", input_dim, latent_dim)
    mlp("model.encoder_sigma", input_dim, latent_dim)
    mlp("model.decoder", latent_dim, input_dim)
    return shapes


def classifier_param_shapes(input_dim: int, num_classes, num_layers: int = 1, hidden_dim: int = 128):
    """Shapes/keys of LatentClassifier's state_dict (latent_classifier.py:30-56)."""
    shapes = []
    single = isinstance(num_classes, int)
    if single:
        if num_layers == 1:
            shapes += [("classifier.layers.0.weight", (num_classes, input_dim)), ("classifier.layers.0.bias", (num_classes,))]
        else:
            shapes += [("classifier.layers.0.weight", (hidden_dim, input_dim)), ("classifier.layers.0.bias", (hidden_dim,))]
            li = 0
            for _ in range(num_layers - 2):
                li += 2
                shapes += [(f"classifier.layers.{li}.weight", (hidden_dim, hidden_dim)), (f"classifier.layers.{li}.bias", (hidden_dim,))]
            li += 2
            shapes += [(f"classifier.layers.{li}.weight", (num_classes, hidden_dim)), (f"classifier.layers.{li}.bias", (num_classes,))]
    else:
        if num_layers == 1:
            feat = input_dim
        else:
            shapes += [("classifier.layers.0.weight", (hidden_dim, input_dim)), ("classifier.layers.0.bias", (hidden_dim,))]
            li = 0
            for _ in range(num_layers - 2):
                li += 2
                shapes += [(f"classifier.layers.{li}.weight", (hidden_dim, hidden_dim)), (f"classifier.layers.{li}.bias", (hidden_dim,))]
            feat = hidden_dim
        for label, c in num_classes.items():
            shapes += [(f"classifier.output_layers.{label}.weight", (c, feat)), (f"classifier.output_layers.{label}.bias", (c,))]
    return shapes


def synth_params(shapes, seed: int = 0, dtype=np.float32) -> Params:
    """nn.Linear-like U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from numpy's PCG64 (portable, versions-stable)."""
    rng = np.random.default_rng(seed)
    out = {}
    fan_in = 1
    for name, shp in shapes:
        if name.endswith(".weight"):
            fan_in = shp[1]
        bound = 1.0 / math.sqrt(fan_in)
        out[name] = rng.uniform(-bound, bound, size=shp).astype(dtype)
    return out


def synth_batch(B: int, D: int, L: int, num_classes=2, seed: int = 1234, dtype=np.float32, unit_norm: bool = True):
    """SURVEY 8(d) synthetic inputs: x ~ N(0,1) rows L2-normalised, y from the CV gender marginals
    (plots/dataset_info_train.json:161-176 through utils.py:92-119), eps ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, D))
    if unit_norm:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    probs = {2: [0.717, 0.283], 3: [0.694, 0.274, 0.032]}
    if isinstance(num_classes, dict):
        y = {k: rng.choice(c, size=B, p=probs.get(c)).astype(np.int64) for k, c in num_classes.items()}
    else:
        y = rng.choice(num_classes, size=B, p=probs.get(num_classes)).astype(np.int64)
    eps = rng.standard_normal((B, L))
    return x.astype(dtype), y, eps.astype(dtype)


# --------------------------------------------------------------------------------------------
# the analysis Langevin variant (analysis/sample_gender_transformation.py:57-99)
# --------------------------------------------------------------------------------------------
def latent_transformation(params: Params, z_start: np.ndarray, noises: Sequence[np.ndarray], target, step_size: float = 0.02, noise_weight: float = 0.0,
                          prior_weight: float = 0.5, threshold: float = 0.95, activation: str = "relu"):
    """Per-sample restatement of the script's loop, batched: every sample ascends log p(y|z) + prior_weight * log p(z) from its start latent
    (the script starts from the encoder mean of a real embedding, :57-60) and stops after the update of the first step whose p(y|z)
    exceeded ``threshold`` (:97-98).  Returns (z_final [N,L], stop_step [N] (len(noises) if never), last evaluated p(y|z) [N])."""
    dt = z_start.dtype.type
    z = z_start.copy()
    N = z.shape[0]
    active = np.ones(N, dtype=bool)
    stop = np.full(N, len(noises), dtype=np.int64)
    prob = np.zeros(N, dtype=z.dtype)
    for step, noise in enumerate(noises):
        g, log_p_y, _ = langevin_grad(params, z, target, activation)          # d/dz [log p(y|z)] - z
        grad = g + z - dt(prior_weight) * z                                     # prior term weighted
        p = np.exp(log_p_y)
        upd = z + dt(0.5 * step_size ** 2) * grad + dt(step_size * noise_weight) * noise
        z = np.where(active[:, None], upd, z)
        prob = np.where(active, p, prob)
        hit = active & (p > threshold)
        stop[hit] = step
        active &= ~hit
        if not active.any():
            break
    return z, stop, prob


# --------------------------------------------------------------------------------------------
# the stand-alone EmbeddingClassifier trainer (ps_vae/embedding_classifier/embedding_classifier.py:64-100)
# --------------------------------------------------------------------------------------------
def embedding_classifier_loss_and_grads(params: Params, x: np.ndarray, y: np.ndarray, compute_grads: bool = True):
    """training_step / validation_step of the reference's EmbeddingClassifier: CrossEntropyLoss(fc3(relu(fc2(relu(fc1(x))))), y) (mean) and
    the multiclass accuracy; closed-form gradients of the six tensors.  Returns (dict(loss, acc), logits, grads | None)."""
    logits, (a1, a2) = embedding_classifier_forward(params, x)
    loss, dlog = cross_entropy(logits, y)
    scal = dict(loss=loss, acc=(logits.argmax(-1) == y).mean())
    if not compute_grads:
        return scal, logits, None
    g = {"fc3.weight": dlog.T @ a2, "fc3.bias": dlog.sum(axis=0)}
    d2 = (dlog @ params["fc3.weight"]) * (a2 > 0)
    g["fc2.weight"] = d2.T @ a1
    g["fc2.bias"] = d2.sum(axis=0)
    d1 = (d2 @ params["fc2.weight"]) * (a1 > 0)
    g["fc1.weight"] = d1.T @ x
    g["fc1.bias"] = d1.sum(axis=0)
    return scal, logits, g
