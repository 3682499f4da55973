"""GPU diagnostics: per-tensor errors of the train step against the numpy oracle (prints, never asserts)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ps_vae_oracle as O  # noqa: E402
from tests.golden_util import GOLDEN, case_batch, case_params, load, rel_err  # noqa: E402
from tests import gpu_util as G  # noqa: E402


def step_report(tag, cfg, B, precision, seed=0, batch=None):
    params = case_params(cfg, np.float64)
    nc = cfg["clf"]["num_classes"] if cfg.get("clf") else 2
    if batch is None:
        x, y, eps = O.synth_batch(B, cfg["D"], cfg["L"], nc, seed=seed)
    else:
        x, y, eps = batch
    scal, out, grads = O.train_loss_and_grads(params, x.astype(np.float64), y, eps.astype(np.float64), kl_loss_weight=cfg.get("kl_w", 1.0),
                                              classifier_loss_weight=cfg.get("clf_w", 1.0), normalize_decoder=cfg.get("normalize_decoder", False),
                                              use_cos_loss=cfg.get("use_cos_loss", False), classifier_activation=(cfg.get("clf") or {}).get("activation", "relu"))
    m = G.module_from_cfg(cfg, precision)
    hot = m.hot_path
    g = torch.empty(hot.arena.numel, device=G.DEV)
    yt = G.labels_to_torch(y) if cfg.get("clf") else None
    losses, _, outs = hot.step(torch.from_numpy(x).to(G.DEV), yt, torch.from_numpy(eps).to(G.DEV), kl_weight=cfg.get("kl_w", 1.0),
                               clf_weight=cfg.get("clf_w", 1.0), use_cos_loss=cfg.get("use_cos_loss", False), grads=g, want_outputs=True)
    torch.cuda.synchronize()
    lt = losses.cpu().numpy()
    print(f"== {tag} B={x.shape[0]} {precision}: loss {lt[0]:.8f}/{float(scal['loss']):.8f} recon {lt[1]:.8f}/{float(scal['recon_loss']):.8f} "
          f"kl {lt[2]:.8f}/{float(scal['kl_loss']):.8f} clf {lt[3]:.8f}/{float(scal['classifier_loss']):.8f}")
    print("   x_hat %.2e mu %.2e ls %.2e" % (rel_err(outs[0].cpu().numpy(), out["x_hat"]), rel_err(outs[1].cpu().numpy(), out["mu"]),
                                             rel_err(outs[2].cpu().numpy(), out["log_sigma"])))
    gd = G.flat_to_dict(m, g)
    print("   grads: " + "  ".join(f"{k.replace('model.', '').replace('.weight', '.w').replace('.bias', '.b')}={rel_err(gd[k], grads[k]):.1e}" for k in grads))


def adam_report():
    L = G.L
    z = np.load(os.path.join(GOLDEN, "adam_cosine.npz"))
    for wd, tag in ((0.0, "wd0"), (0.01, "wd01")):
        p = torch.from_numpy(z["p0"].copy()).to(G.DEV)
        m = torch.zeros_like(p)
        v = torch.zeros_like(p)
        shadow = torch.empty(p.numel(), dtype=torch.bfloat16, device=G.DEV)
        for s, g in enumerate(z["grads"]):
            gt = torch.from_numpy(g.astype(np.float32)).to(G.DEV)
            L.check(L.lib().psvae_adam_step(p.data_ptr(), gt.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 3e-3, 0.9, 0.999, 1e-8, wd, s + 1,
                                            1.0, shadow.data_ptr(), G.stream()))
            ref = z[f"{tag}/f32/p{s}"]
            ref64 = z[f"{tag}/f64/p{s}"]
            got = p.cpu().numpy()
            print(f"adam {tag} step {s}: max|p-ref32|/max|ref| = {np.abs(got - ref).max() / np.abs(ref).max():.2e}  vs f64 {np.abs(got - ref64).max() / np.abs(ref64).max():.2e} "
                  f"(ref32 vs f64 {np.abs(ref - ref64).max() / np.abs(ref64).max():.2e})  shadow_equal={torch.equal(shadow, p.to(torch.bfloat16))}")
        print("   m %.2e v %.2e" % (rel_err(m.cpu().numpy(), z[f"{tag}/f32/m"]), rel_err(v.cpu().numpy(), z[f"{tag}/f32/v"])))


if __name__ == "__main__":
    which = sys.argv[1:] or ["adam", "normcos", "ragged", "big", "golden_bf16"]
    if "adam" in which:
        adam_report()
    if "normcos" in which:
        z, cfg = load("train_d256_norm_cos")
        for prec in ("fp32", "bf16"):
            step_report("norm_cos", cfg, cfg["B"], prec, batch=case_batch(cfg, 0, np.float32))
        cfg2 = dict(cfg); cfg2["use_cos_loss"] = False
        step_report("norm_only", cfg2, 16, "fp32", batch=case_batch(cfg, 0, np.float32))
        cfg3 = dict(cfg); cfg3["normalize_decoder"] = False
        step_report("cos_only", cfg3, 16, "fp32", batch=case_batch(cfg, 0, np.float32))
    if "ragged" in which:
        cfg = dict(D=256, L=64, wseed=3, clf=dict(input_dim=64, num_classes=3))
        for B in (129, 512, 1000, 1024, 4096):
            step_report("ragged", cfg, B, "fp32", seed=B)
    if "golden_bf16" in which:
        for name in ("train_d256_c2", "train_d192_noclf", "train_d512_c3_mlp"):
            z, cfg = load(name)
            step_report(name, cfg, cfg["B"], "bf16", batch=case_batch(cfg, 0, np.float32))
        cfg = dict(D=256, L=64, wseed=3, clf=dict(input_dim=64, num_classes=2))
        for B in (1024, 8192):
            step_report("bf16", cfg, B, "bf16", seed=B)
    if "big" in which:
        cfg = dict(D=256, L=64, wseed=3, clf=dict(input_dim=64, num_classes=2))
        step_report("big", cfg, 65536, "fp32", seed=1234)
