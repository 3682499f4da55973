"""Diagnostic: deterministic-mode step with and without programmatic dependent launch -- which outputs differ?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ps_vae_oracle as O
from tests import gpu_util as G
cfg = dict(D=256, L=64, wseed=3, clf=dict(input_dim=64, num_classes=2))
module = G.module_from_cfg(cfg, "bf16")
hot = module.hot_path
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192 + 77
x, y, eps = O.synth_batch(B, 256, 64, 2, seed=78)
xt, yt, et = torch.from_numpy(x).to(G.DEV), torch.from_numpy(y).to(G.DEV), torch.from_numpy(eps).to(G.DEV)
def run():
    g = torch.empty(hot.arena.numel, device=G.DEV)
    losses, _, outs = hot.step(xt, yt, et, grads=g, want_outputs=True)
    torch.cuda.synchronize()
    return g.clone(), losses.clone(), [o.clone() for o in outs]
G.L.set_option("deterministic", 1)
res = {}
for pdl in (1, 0, 1, 0):
    G.L.set_option("pdl", pdl)
    res.setdefault(pdl, []).append(run())
G.L.set_option("pdl", 1)
for a, b, name in ((res[1][0], res[1][1], "pdl=1 vs pdl=1"), (res[0][0], res[0][1], "pdl=0 vs pdl=0"), (res[1][0], res[0][0], "pdl=1 vs pdl=0")):
    print(name, "losses equal", torch.equal(a[1], b[1]), "grads equal", torch.equal(a[0], b[0]), "outs equal", all(torch.equal(p, q) for p, q in zip(a[2], b[2])))
    if not torch.equal(a[1], b[1]):
        print("  losses", a[1].cpu().numpy()[:10], b[1].cpu().numpy()[:10])
    if not torch.equal(a[0], b[0]):
        da, db = G.flat_to_dict(module, a[0]), G.flat_to_dict(module, b[0])
        for k in da:
            if not np.array_equal(da[k], db[k]):
                d = np.abs(da[k].astype(np.float64) - db[k]).max()
                print(f"  {k}: max abs diff {d:.3e} (max abs {np.abs(da[k]).max():.3e}), {np.sum(da[k] != db[k])} of {da[k].size} differ")
