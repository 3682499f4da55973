"""Helpers shared by the CPU (oracle-vs-golden) and GPU (kernel-vs-oracle/golden) parity tests."""
import json
import os

import numpy as np

from oracle import ps_vae_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_SAMPLE = 64


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg"])) if "cfg" in z.files else None
    return z, cfg


def sample_idx(name: str, size: int) -> np.ndarray:
    # must mirror oracle/make_golden.py::sample_idx
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31)
    rng = np.random.default_rng(seed)
    return rng.integers(0, size, size=min(N_SAMPLE, size))


def case_params(cfg, dtype=np.float32):
    shapes = O.vae_param_shapes(cfg["D"], cfg["L"])
    if cfg.get("clf"):
        c = cfg["clf"]
        shapes += O.classifier_param_shapes(c["input_dim"], c["num_classes"], c.get("num_layers", 1), c.get("hidden_dim", 128))
    p = O.synth_params(shapes, seed=cfg["wseed"], dtype=np.float64)
    return {k: v.astype(np.float32).astype(dtype) for k, v in p.items()}


def case_consistency_params(cfg, dtype=np.float32):
    """Synthetic weights of the frozen consistency classifier of a golden case (None when the case has none)."""
    c = cfg.get("cons")
    if not c:
        return None
    p = O.synth_params(O.embedding_classifier_param_shapes(cfg["D"], c["num_classes"], c.get("hidden_dim", 128)), seed=c["wseed"], dtype=np.float64)
    return {k: v.astype(np.float32).astype(dtype) for k, v in p.items()}


def case_batch(cfg, step, dtype=np.float32):
    nc = cfg["clf"]["num_classes"] if cfg.get("clf") else 2
    x, y, eps = O.synth_batch(cfg["B"], cfg["D"], cfg["L"], nc, seed=cfg["dseed"] + step)
    return x.astype(dtype), y, eps.astype(dtype)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.sqrt((b * b).sum())
    return np.sqrt(((a - b) ** 2).sum()) / (den if den > 0 else 1.0)


def check_summary(z, tag, name, arr, tol):
    """Compare a big tensor with its golden (sum, l2, sampled entries); returns the worst relative error."""
    a = np.asarray(arr, dtype=np.float64).ravel()
    l2 = float(z[f"{tag}/{name}/l2"])
    samples = z[f"{tag}/{name}/samples"]
    got = a[sample_idx(name, a.size)]
    # ||a-b||/||b|| estimated on the sampled entries (the per-tensor bar of north_star), plus two
    # whole-tensor aggregates: the l2 norm, and the plain sum (|sum(a-b)| <= sqrt(n)*||a-b||: a coarse 16x slack catches sign/offset bugs)
    e_samples = np.sqrt(((got - samples) ** 2).sum()) / (np.sqrt((samples ** 2).sum()) + 1e-30)
    e_l2 = abs(np.sqrt((a * a).sum()) - l2) / (l2 + 1e-30)
    e_sum = abs(a.sum() - float(z[f"{tag}/{name}/sum"])) / (l2 + 1e-30) / 16
    worst = max(e_samples, e_l2, e_sum)
    assert worst <= tol, f"{tag}/{name}: samples {e_samples:.3e} l2 {e_l2:.3e} sum {e_sum:.3e} > {tol}"
    return worst
