"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this container.

    python -m oracle.make_golden            # rewrites tests/golden/

Weights/inputs come from oracle.ps_vae_oracle.synth_params / synth_batch (numpy PCG64, reproducible
anywhere) and are loaded into the reference's own modules with load_state_dict, so the fixtures only
need to store seeds plus the reference's outputs.  Large tensors (gradients, post-Adam parameters) are
stored as (sum, l2 norm, 64 sampled entries) per tensor; small ones in full.  Every value is recorded
twice: from the fp32 reference at torch.set_float32_matmul_precision('highest') and from its
``.double()`` twin.
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ps_vae_oracle as O  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, injected_normals, load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
N_SAMPLE = 64


def sample_idx(name: str, size: int) -> np.ndarray:
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31)
    rng = np.random.default_rng(seed)
    return rng.integers(0, size, size=min(N_SAMPLE, size))


def summarize(store: dict, tag: str, name: str, arr: np.ndarray):
    a = np.asarray(arr, dtype=np.float64).ravel()
    store[f"{tag}/{name}/sum"] = a.sum()
    store[f"{tag}/{name}/l2"] = np.sqrt((a * a).sum())
    store[f"{tag}/{name}/samples"] = a[sample_idx(name, a.size)]


def build_module(ref, cfg, dtype):
    hp = dict(model=dict(input_dim=cfg["D"], latent_dim=cfg["L"], normalize_decoder=cfg.get("normalize_decoder", False)),
              optimizer=cfg.get("optimizer", dict(lr=1e-3)), scheduler=dict(T_max=200),
              kl_loss_weight=cfg.get("kl_w", 1.0), classifier_loss_weight=cfg.get("clf_w", 1.0),
              use_cos_loss=cfg.get("use_cos_loss", False), consistency_loss_weight=cfg.get("cons_w", 1.0))
    if cfg.get("clf"):
        hp["classifier"] = dict(cfg["clf"])
    if isinstance(hp.get("classifier", {}).get("num_classes"), dict):
        # the Lightning multi-label branch is broken in the reference: build the pieces by hand
        clf_h = hp.pop("classifier")
        m = ref.PseudoSpeakerVAE(**hp)
        m.classifier = ref.LatentClassifier(**clf_h)
    else:
        m = ref.PseudoSpeakerVAE(**hp)
    shapes = O.vae_param_shapes(cfg["D"], cfg["L"])
    if cfg.get("clf"):
        c = cfg["clf"]
        shapes += O.classifier_param_shapes(c["input_dim"], c["num_classes"], c.get("num_layers", 1), c.get("hidden_dim", 128))
    params = O.synth_params(shapes, seed=cfg["wseed"], dtype=np.float64)
    sd = {k: torch.from_numpy(v.astype(np.float32)) for k, v in params.items()}
    missing = m.load_state_dict(sd, strict=True)
    if cfg.get("cons"):
        # lightning.py:44-52 loads a frozen EmbeddingClassifier from a Lightning checkpoint; the stub has no checkpoint reader, so the
        # (unmodified) reference class is built here, given synthetic weights, frozen and put in eval mode exactly as those lines do
        import importlib

        EC = importlib.import_module("ps_vae.embedding_classifier.embedding_classifier").EmbeddingClassifier
        c = cfg["cons"]
        ec = EC(input_dim=cfg["D"], num_classes=c["num_classes"], hidden_dim=c.get("hidden_dim", 128))
        cp = O.synth_params(O.embedding_classifier_param_shapes(cfg["D"], c["num_classes"], c.get("hidden_dim", 128)), seed=c["wseed"], dtype=np.float64)
        ec.load_state_dict({k: torch.from_numpy(v.astype(np.float32)) for k, v in cp.items()}, strict=False)
        for p in ec.parameters():
            p.requires_grad = False
        ec.eval()
        m.consistency_classifier = ec
    m = m.to(dtype)
    if dtype == torch.float64:
        # fp64 twin holds the same fp32-rounded values
        pass
    return m


def train_case(ref, name, cfg):
    store = {"cfg": json.dumps(cfg)}
    steps = cfg.get("steps", 3)
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = build_module(ref, cfg, dtype)
        opt = m.configure_optimizers()["optimizer"]
        for s in range(steps):
            x, y, eps = O.synth_batch(cfg["B"], cfg["D"], cfg["L"], cfg["clf"]["num_classes"] if cfg.get("clf") else 2, seed=cfg["dseed"] + s)
            xt = torch.from_numpy(x).to(dtype)
            yt = torch.from_numpy(y)
            et = torch.from_numpy(eps).to(dtype)
            opt.zero_grad()
            with injected_normals([et]):
                x_hat, mu, ls = m(xt)
            with injected_normals([et]):
                loss = m.training_step((xt, yt), 0)["loss"]
            loss.backward()
            st = f"{tag}/step{s}"
            for k in ("x_hat", "mu", "ls"):
                store[f"{st}/{k}"] = {"x_hat": x_hat, "mu": mu, "ls": ls}[k].detach().numpy()
            for k, v in m.logged.items():
                store[f"{st}/log/{k}"] = float(v)
            for pn, p in m.named_parameters():
                if p.grad is None:      # the frozen consistency classifier
                    continue
                summarize(store, f"{st}/grad", pn, p.grad.detach().numpy())
            opt.step()
            for pn, p in m.named_parameters():
                if pn.startswith("consistency_classifier."):
                    continue
                summarize(store, f"{st}/param", pn, p.detach().numpy())
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **store)
    print("wrote", name, len(store))


def sampling_case(ref, name, cfg):
    store = {"cfg": json.dumps(cfg)}
    N, L = cfg["N"], cfg["L"]
    rng = np.random.default_rng(cfg["dseed"])
    z0 = rng.standard_normal((N, L)).astype(np.float32)
    noises = [rng.standard_normal((N, L)).astype(np.float32) for _ in range(cfg["steps"])]
    store["z0"] = z0
    store["noises"] = np.stack(noises) if noises else np.zeros((0, N, L), np.float32)
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = build_module(ref, cfg, dtype)
        torch.set_default_dtype(dtype)
        try:
            with injected_normals([torch.from_numpy(z0).to(dtype)]):
                xu = ref.unconditional_synthesis(m, N, "cpu")
            store[f"{tag}/uncond"] = xu.numpy()
            if cfg.get("clf"):
                inj = [torch.from_numpy(z0).to(dtype)] + [torch.from_numpy(n).to(dtype) for n in noises]
                with injected_normals(inj):
                    xc, hist = ref.conditional_synthesis(m, N, cfg["target"], step_size=cfg["step_size"], num_steps=cfg["steps"],
                                                         noise_weight=cfg["noise_weight"], return_history=True, device="cpu")
                store[f"{tag}/cond"] = xc.numpy()
                store[f"{tag}/hist"] = np.stack(hist)
        finally:
            torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **store)
    print("wrote", name, len(store))


def transformation_case(ref, name, cfg):
    """The Langevin variant of analysis/sample_gender_transformation.py:57-99, run per sample exactly as that script does (batch of one,
    torch autograd through the reference's own modules): start from the encoded embedding (`_, z, _ = vae_model(embed)`: z = mu),
    ascend log p(y|z) + PRIOR_WEIGHT * log p(z) with step 0.5 * STEP_SIZE^2, optional noise, stop after the update of the first step
    whose p(y|z) exceeded THRESHOLD.  Stored: the inputs, the start latents, the final latents, the step at which each sample stopped
    (max_steps if it never did) and the classifier probability of its last evaluated step."""
    import torch.nn.functional as F

    store = {"cfg": json.dumps(cfg)}
    N, L, D = cfg["N"], cfg["L"], cfg["D"]
    x, _, _ = O.synth_batch(N, D, L, 2, seed=cfg["dseed"])
    rng = np.random.default_rng(cfg["dseed"] + 1)
    noises = rng.standard_normal((cfg["max_steps"], N, L)).astype(np.float32)
    store["x"] = x
    store["noises"] = noises
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        vae_model = build_module(ref, cfg, dtype)
        vae_model.eval()
        z_start, z_final, stops, probs = [], [], [], []
        for i in range(N):
            emebd = torch.from_numpy(x[i]).to(dtype)
            with torch.no_grad():
                _, z, _ = vae_model(emebd.unsqueeze(0))
                z.requires_grad_()
            z_start.append(z.detach().numpy()[0].copy())
            classifier_target = cfg["target"]
            stop = cfg["max_steps"]
            for step in range(cfg["max_steps"]):
                logits = vae_model.classifier(z)
                log_probs = F.log_softmax(logits, dim=-1)
                log_p_y_given_z = log_probs[:, classifier_target]
                log_p_z = -0.5 * (z ** 2).sum(dim=1)
                log_p_z_given_y = log_p_y_given_z + cfg["prior_weight"] * log_p_z
                grad = torch.autograd.grad(log_p_z_given_y.sum(), z)[0]
                noise = torch.from_numpy(noises[step, i:i + 1]).to(dtype)
                p_y_given_z = log_p_y_given_z.clone().exp()
                z = z + 0.5 * (cfg["step_size"] ** 2) * grad + cfg["step_size"] * cfg["noise_weight"] * noise
                z.requires_grad_()
                if p_y_given_z.item() > cfg["threshold"]:
                    stop = step
                    break
            z_final.append(z.detach().numpy()[0].copy())
            stops.append(stop)
            probs.append(float(p_y_given_z.item()))
        store[f"{tag}/z_start"] = np.stack(z_start)
        store[f"{tag}/z_final"] = np.stack(z_final)
        store[f"{tag}/stop"] = np.array(stops, dtype=np.int64)
        store[f"{tag}/prob"] = np.array(probs, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **store)
    print("wrote", name, "stops", store["f64/stop"].tolist())


def embedding_classifier_case(name, cfg):
    """The stand-alone EmbeddingClassifier trainer (ps_vae/embedding_classifier/embedding_classifier.py:38-100): the unmodified module's
    training_step -> backward -> its own configure_optimizers() Adam, a validation_step on the next batch; logits, logged metrics, every
    gradient and the post-step parameters."""
    import importlib

    load_reference()
    EC = importlib.import_module("ps_vae.embedding_classifier.embedding_classifier").EmbeddingClassifier
    store = {"cfg": json.dumps(cfg)}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        m = EC(cfg["D"], cfg["num_classes"], cfg["hidden_dim"], optimizer_cfg=dict(cfg["optimizer"]))
        cp = O.synth_params(O.embedding_classifier_param_shapes(cfg["D"], cfg["num_classes"], cfg["hidden_dim"]), seed=cfg["wseed"], dtype=np.float64)
        m.load_state_dict({k: torch.from_numpy(v.astype(np.float32)) for k, v in cp.items()})
        m = m.to(dtype)
        opt = m.configure_optimizers()
        for s in range(cfg["steps"]):
            x, y, _ = O.synth_batch(cfg["B"], cfg["D"], 64, cfg["num_classes"], seed=cfg["dseed"] + s)
            xt, yt = torch.from_numpy(x).to(dtype), torch.from_numpy(y)
            opt.zero_grad()
            loss = m.training_step((xt, yt), s)
            loss.backward()
            st = f"{tag}/step{s}"
            store[f"{st}/logits"] = m(xt).detach().numpy()
            for k, v in m.logged.items():
                store[f"{st}/log/{k}"] = float(v)
            for pn, p in m.named_parameters():
                store[f"{st}/grad/{pn}"] = p.grad.detach().numpy().copy()
            opt.step()
            for pn, p in m.named_parameters():
                store[f"{st}/param/{pn}"] = p.detach().numpy().copy()
        x, y, _ = O.synth_batch(cfg["B"], cfg["D"], 64, cfg["num_classes"], seed=cfg["dseed"] + 99)
        m.validation_step((torch.from_numpy(x).to(dtype), torch.from_numpy(y)), 0)
        for k in ("val_acc", "val_loss"):
            store[f"{tag}/val/{k}"] = float(m.logged[k])
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **store)
    print("wrote", name, len(store))


def adam_case():
    store = {}
    rng = np.random.default_rng(77)
    p0 = rng.standard_normal(1000).astype(np.float32)
    gs = [rng.standard_normal(1000).astype(np.float32) * 10.0 ** rng.integers(-4, 1) for _ in range(4)]
    store["p0"] = p0
    store["grads"] = np.stack(gs)
    for wd, tag in ((0.0, "wd0"), (0.01, "wd01")):
        for dtype, dt in ((torch.float32, "f32"), (torch.float64, "f64")):
            p = torch.nn.Parameter(torch.from_numpy(p0.copy()).to(dtype))
            opt = torch.optim.Adam([p], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, foreach=False)
            for s, g in enumerate(gs):
                p.grad = torch.from_numpy(g.copy()).to(dtype)
                opt.step()
                store[f"{tag}/{dt}/p{s}"] = p.detach().numpy().copy()
            st = opt.state[p]
            store[f"{tag}/{dt}/m"] = st["exp_avg"].numpy()
            store[f"{tag}/{dt}/v"] = st["exp_avg_sq"].numpy()
    # cosine schedule (lightning.py:206-213 steps it once per epoch)
    for T_max, eta_min, n in ((200, 0.0, 6), (3, 1e-5, 9)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=1e-3)
        sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=T_max, eta_min=eta_min)
        lrs = [opt.param_groups[0]["lr"]]
        for _ in range(n):
            opt.step()
            sch.step()
            lrs.append(opt.param_groups[0]["lr"])
        store[f"cosine/T{T_max}"] = np.array(lrs, dtype=np.float64)
        store[f"cosine/T{T_max}/eta_min"] = eta_min
    np.savez_compressed(os.path.join(OUT, "adam_cosine.npz"), **store)
    print("wrote adam_cosine")


def label_tables():
    """ps_vae/utils.py cannot be imported here (matplotlib/sklearn/torchvision absent): read the dict
    literals of the three map_* functions with ``ast`` instead, and evaluate them the way the functions do."""
    src = open(os.path.join(REFERENCE_ROOT, "ps_vae", "utils.py")).read()
    tree = ast.parse(src)
    tables = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name.startswith("map_"):
            for sub in ast.walk(node):
                if isinstance(sub, ast.Assign) and isinstance(sub.value, ast.Dict):
                    tables[node.name] = ast.literal_eval(sub.value)
            default = [ast.literal_eval(s.value.args[1]) for s in ast.walk(node) if isinstance(s, ast.Return) and isinstance(s.value, ast.Call)]
            tables[node.name + "::default"] = default[0]
    # target parsing (inference.py:128-132) evaluated on a few CLI strings
    parsed = {}
    for text in ["1", "0", "2", '{"age": 2, "gender": 1}', "{\"gender\": 0}"]:
        try:
            parsed[text] = json.loads(text)
        except json.JSONDecodeError:
            parsed[text] = int(text)
    with open(os.path.join(OUT, "label_tables.json"), "w") as f:
        json.dump({"tables": tables, "parsed_targets": parsed, "sample_name_fmt": "sample_{i}.pt"}, f, indent=1, sort_keys=True)
    print("wrote label_tables", list(tables))


def main():
    torch.set_float32_matmul_precision("highest")
    torch.manual_seed(0)
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    if "--only-embclf" in sys.argv:
        embedding_classifier_case("embclf_d256_c3", dict(D=256, num_classes=3, hidden_dim=128, B=48, wseed=51, dseed=1300, steps=2,
                                                          optimizer=dict(lr=2e-3, weight_decay=0.01)))
        embedding_classifier_case("embclf_d192_c2", dict(D=192, num_classes=2, hidden_dim=64, B=33, wseed=52, dseed=1400, steps=2, optimizer=dict()))
        return
    if "--only-transform" in sys.argv:
        transformation_case(ref, "transform_prior_threshold", dict(D=256, L=64, N=12, wseed=41, dseed=1100, max_steps=40, step_size=0.5, noise_weight=0.0,
                                                                   prior_weight=0.5, threshold=0.64, target=1, clf=dict(input_dim=64, num_classes=2)))
        transformation_case(ref, "transform_noise_c3", dict(D=192, L=64, N=10, wseed=42, dseed=1200, max_steps=30, step_size=0.9, noise_weight=0.2,
                                                            prior_weight=0.25, threshold=0.5, target=2, clf=dict(input_dim=64, num_classes=3)))
        return
    # consistency-classifier cases (lightning.py:44-52,100-108), added after the first fixture set: `--only-cons` writes just these
    train_case(ref, "train_d256_c2_cons", dict(D=256, L=64, B=24, wseed=15, dseed=910, steps=2, cons_w=0.7,
                                               clf=dict(input_dim=64, num_classes=2), cons=dict(num_classes=2, hidden_dim=128, wseed=31)))
    train_case(ref, "train_d192_c3_cons_norm_cos", dict(D=192, L=64, B=16, wseed=16, dseed=950, steps=2, cons_w=1.5, normalize_decoder=True,
                                                        use_cos_loss=True, clf=dict(input_dim=64, num_classes=3),
                                                        cons=dict(num_classes=3, hidden_dim=64, wseed=32)))
    if "--only-cons" in sys.argv:
        return
    # the analysis Langevin variant (analysis/sample_gender_transformation.py:57-99), added in round 2: `--only-transform` writes just these
    transformation_case(ref, "transform_prior_threshold", dict(D=256, L=64, N=12, wseed=41, dseed=1100, max_steps=40, step_size=0.5, noise_weight=0.0,
                                                               prior_weight=0.5, threshold=0.64, target=1, clf=dict(input_dim=64, num_classes=2)))
    transformation_case(ref, "transform_noise_c3", dict(D=192, L=64, N=10, wseed=42, dseed=1200, max_steps=30, step_size=0.9, noise_weight=0.2,
                                                        prior_weight=0.25, threshold=0.5, target=2, clf=dict(input_dim=64, num_classes=3)))
    if "--only-transform" in sys.argv:
        return
    embedding_classifier_case("embclf_d256_c3", dict(D=256, num_classes=3, hidden_dim=128, B=48, wseed=51, dseed=1300, steps=2,
                                                      optimizer=dict(lr=2e-3, weight_decay=0.01)))
    embedding_classifier_case("embclf_d192_c2", dict(D=192, num_classes=2, hidden_dim=64, B=33, wseed=52, dseed=1400, steps=2, optimizer=dict()))
    train_case(ref, "train_d256_c2", dict(D=256, L=64, B=32, wseed=11, dseed=100, clf=dict(input_dim=64, num_classes=2)))
    train_case(ref, "train_d192_noclf", dict(D=192, L=64, B=16, wseed=12, dseed=200, kl_w=0.5, steps=2))
    train_case(ref, "train_d512_c3_mlp", dict(D=512, L=64, B=24, wseed=13, dseed=300, clf_w=2.0, steps=2,
                                              clf=dict(input_dim=64, num_classes=3, num_layers=2, hidden_dim=128)))
    train_case(ref, "train_d256_norm_cos", dict(D=256, L=64, B=16, wseed=14, dseed=400, steps=2, normalize_decoder=True, use_cos_loss=True,
                                                clf=dict(input_dim=64, num_classes=2), optimizer=dict(lr=2e-3, weight_decay=0.01)))
    sampling_case(ref, "sample_single", dict(D=256, L=64, N=16, wseed=21, dseed=500, steps=6, step_size=0.05, noise_weight=1.0, target=1,
                                             clf=dict(input_dim=64, num_classes=2)))
    sampling_case(ref, "sample_c3_mlp", dict(D=256, L=64, N=12, wseed=22, dseed=600, steps=5, step_size=0.1, noise_weight=0.5, target=2,
                                             clf=dict(input_dim=64, num_classes=3, num_layers=3, hidden_dim=128, activation="tanh")))
    sampling_case(ref, "sample_multilabel", dict(D=192, L=64, N=10, wseed=23, dseed=700, steps=4, step_size=0.1, noise_weight=1.0,
                                                 target={"age": 2, "gender": 1},
                                                 clf=dict(input_dim=64, num_classes={"age": 3, "gender": 2}, num_layers=2, hidden_dim=128)))
    sampling_case(ref, "sample_multilabel_1layer", dict(D=256, L=64, N=8, wseed=24, dseed=800, steps=3, step_size=0.2, noise_weight=1.0,
                                                        target={"gender": 0, "age": 1}, normalize_decoder=True,
                                                        clf=dict(input_dim=64, num_classes={"age": 3, "gender": 2})))
    adam_case()
    label_tables()


if __name__ == "__main__":
    main()
