"""Per-tensor errors of the bf16 tensor-core step against the bf16-emulating twin for a few batch sizes (diagnosis of a failing
tests/test_gpu_round2.py case).  python tools/twin_diag.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bf16_twin as T, ps_vae_oracle as O
from tests import gpu_util as G
from tests.golden_util import case_params, rel_err

cfg = dict(D=256, L=64, wseed=3, clf=dict(input_dim=64, num_classes=2))
params = case_params(cfg, np.float32)
for det in (0, 1):
    G.L.set_option("deterministic", det)
    for B in (1, 2, 3, 4, 5, 8, 129):
        x, y, eps = O.synth_batch(B, 256, 64, 2, seed=B + 5)
        scal, out, grads = T.train_loss_and_grads_bf16(params, x, y, eps)
        m = G.module_from_cfg(cfg, "bf16")
        g = torch.empty(m.hot_path.arena.numel, device=G.DEV)
        losses, _, outs = m.hot_path.step(torch.from_numpy(x).to(G.DEV), torch.from_numpy(y).to(G.DEV), torch.from_numpy(eps).to(G.DEV), grads=g, want_outputs=True)
        gd = G.flat_to_dict(m, g)
        errs = {k.replace("model.", ""): rel_err(gd[k], grads[k]) for k in grads}
        f = {k: rel_err(o.cpu().numpy(), out[k]) for o, k in zip(outs, ("x_hat", "mu", "log_sigma"))}
        print(f"det={det} B={B}: loss {float(losses[0]):.7f} vs {float(scal['loss']):.7f} fwd", {k: f"{v:.1e}" for k, v in f.items()},
              "worst", max(errs, key=errs.get), f"{max(errs.values()):.1e}", {k: f"{v:.0e}" for k, v in errs.items() if v > 1e-3}, flush=True)
G.L.set_option("deterministic", 0)
