"""Pins oracle/ (the numpy restatement) to the fixtures the unmodified reference produced
(oracle/make_golden.py), and -- when /root/reference is present -- to the live reference."""
import json
import os

import numpy as np
import pytest

from oracle import philox_ref as P
from oracle import ps_vae_oracle as O
from tests.golden_util import GOLDEN, case_batch, case_consistency_params, case_params, check_summary, load, rel_err

TRAIN_CASES = ["train_d256_c2", "train_d192_noclf", "train_d512_c3_mlp", "train_d256_norm_cos", "train_d256_c2_cons",
               "train_d192_c3_cons_norm_cos"]
SAMPLE_CASES = ["sample_single", "sample_c3_mlp", "sample_multilabel", "sample_multilabel_1layer"]


def _run_train(cfg, z, dtype, tag, tol_out, tol_grad):
    params = case_params(cfg, dtype)
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(p) for k, p in params.items()}
    opt = cfg.get("optimizer", dict(lr=1e-3))
    act = (cfg.get("clf") or {}).get("activation", "relu")
    cons = case_consistency_params(cfg, dtype)
    for s in range(cfg.get("steps", 3)):
        x, y, eps = case_batch(cfg, s, dtype)
        scal, out, grads = O.train_loss_and_grads(
            params, x, y, eps, kl_loss_weight=cfg.get("kl_w", 1.0), classifier_loss_weight=cfg.get("clf_w", 1.0),
            normalize_decoder=cfg.get("normalize_decoder", False), use_cos_loss=cfg.get("use_cos_loss", False),
            classifier_activation=act, consistency_params=cons, consistency_loss_weight=cfg.get("cons_w", 1.0))
        st = f"{tag}/step{s}"
        assert rel_err(out["x_hat"], z[f"{st}/x_hat"]) <= tol_out
        assert rel_err(out["mu"], z[f"{st}/mu"]) <= tol_out
        assert rel_err(out["log_sigma"], z[f"{st}/ls"]) <= tol_out
        for ours, theirs in (("loss", "train_loss"), ("recon_loss", "train_recon_loss"), ("kl_loss", "train_kl_loss")):
            ref = float(z[f"{st}/log/{theirs}"])
            assert abs(float(scal[ours]) - ref) <= tol_out * max(1.0, abs(ref)), (ours, float(scal[ours]), ref)
        if cfg.get("clf"):
            assert abs(float(scal["classifier_loss"]) - float(z[f"{st}/log/train_classifier_loss"])) <= tol_out
            assert abs(float(scal["classifier_acc"]) - float(z[f"{st}/log/train_classifier_acc"])) <= 1e-7
        if cons is not None:
            assert abs(float(scal["consistency_loss"]) - float(z[f"{st}/log/train_consistency_loss"])) <= tol_out
            assert abs(float(scal["consistency_acc"]) - float(z[f"{st}/log/train_consistency"])) <= 1e-7
        for k in params:
            check_summary(z, f"{st}/grad", k, grads[k], tol_grad)
        for k in params:
            params[k], m[k], v[k] = O.adam_step(params[k], grads[k].astype(dtype), m[k], v[k], s + 1, lr=opt.get("lr", 1e-3),
                                                weight_decay=opt.get("weight_decay", 0.0))
            check_summary(z, f"{st}/param", k, params[k], tol_grad)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_step_fp64_twin(name):
    z, cfg = load(name)
    _run_train(cfg, z, np.float64, "f64", 1e-12, 1e-10)


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_train_step_fp32(name):
    """fp32 oracle vs fp32 reference at 'highest': the 1e-5 relative bar of north_star."""
    z, cfg = load(name)
    _run_train(cfg, z, np.float32, "f32", 1e-5, 1e-5)


@pytest.mark.parametrize("name", SAMPLE_CASES)
@pytest.mark.parametrize("dtype,tag,tol", [(np.float64, "f64", 1e-12), (np.float32, "f32", 1e-5)])
def test_sampling(name, dtype, tag, tol):
    z, cfg = load(name)
    params = case_params(cfg, dtype)
    z0 = z["z0"].astype(dtype)
    noises = [n.astype(dtype) for n in z["noises"]]
    nd = cfg.get("normalize_decoder", False)
    act = cfg["clf"].get("activation", "relu")
    xu = O.unconditional_synthesis(params, z0, nd)
    assert rel_err(xu, z[f"{tag}/uncond"]) <= tol
    xc, hist = O.conditional_synthesis(params, z0, noises, cfg["target"], cfg["step_size"], cfg["noise_weight"], nd, act, return_history=True)
    assert rel_err(xc, z[f"{tag}/cond"]) <= tol
    assert rel_err(np.stack(hist), z[f"{tag}/hist"]) <= tol


def test_adam_and_cosine():
    z = np.load(os.path.join(GOLDEN, "adam_cosine.npz"))
    for wd, tag in ((0.0, "wd0"), (0.01, "wd01")):
        for dtype, dt, tol in ((np.float32, "f32", 3e-7), (np.float64, "f64", 1e-14)):
            p = z["p0"].astype(dtype)
            m = np.zeros_like(p)
            v = np.zeros_like(p)
            for s, g in enumerate(z["grads"]):
                p, m, v = O.adam_step(p, g.astype(dtype), m, v, s + 1, lr=3e-3, weight_decay=wd)
                ref = z[f"{tag}/{dt}/p{s}"]
                assert np.abs(p - ref).max() <= tol * np.abs(ref).max(), (tag, dt, s)
            assert rel_err(m, z[f"{tag}/{dt}/m"]) <= 10 * tol
            assert rel_err(v, z[f"{tag}/{dt}/v"]) <= 10 * tol
    for T_max in (200, 3):
        ref = z[f"cosine/T{T_max}"]
        got = O.cosine_annealing_lr(1e-3, T_max, float(z[f"cosine/T{T_max}/eta_min"]), len(ref) - 1)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-18)


def test_label_tables_bit_exact():
    with open(os.path.join(GOLDEN, "label_tables.json")) as f:
        g = json.load(f)
    t = g["tables"]
    assert O.CV_AGE_TO_LABEL == t["map_cv_age_to_label"]
    assert O.CV_GENDER_TO_LABEL == t["map_cv_gender_to_label"]
    assert O.VCTK_GENDER_TO_LABEL == t["map_vctk_gender_to_label"]
    for fn, key in ((O.map_cv_age_to_label, "map_cv_age_to_label"), (O.map_cv_gender_to_label, "map_cv_gender_to_label"),
                    (O.map_vctk_gender_to_label, "map_vctk_gender_to_label")):
        assert fn("no-such-key") == t[key + "::default"] == -1
        assert fn(None) == -1
        for k, v in t[key].items():
            assert fn(k) == v
    for text, val in g["parsed_targets"].items():
        assert O.parse_classifier_target(text) == val
    assert O.sample_filename(7) == g["sample_name_fmt"].format(i=7)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def run(c, k):
        return tuple(int(x) for x in P.philox4x32_10(*[np.uint32(v) for v in c], k[0], k[1]))
    assert run((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert run((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert run((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_philox_normal_sharding_and_moments():
    full = P.philox_normal(64, 64, seed=99, offset=5)
    part = P.philox_normal(16, 64, seed=99, offset=5, row0=32)
    assert np.array_equal(full[32:48], part)          # counter = global element index
    big = P.philox_normal(20000, 64, seed=1, offset=0)
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1) < 5e-3
    assert not np.array_equal(P.philox_normal(4, 64, 1, 0), P.philox_normal(4, 64, 1, 1))


def test_live_reference_agrees_when_present():
    """Belt and braces: run the unmodified reference right here and compare (skipped on the GPU box)."""
    from oracle.ref_loader import reference_available

    if not reference_available():
        pytest.skip("/root/reference not present")
    import torch
    from oracle.ref_loader import injected_normals, load_reference

    torch.set_float32_matmul_precision("highest")
    ref = load_reference()
    cfg = dict(D=256, L=64, B=20, wseed=5, dseed=9, clf=dict(input_dim=64, num_classes=3))
    params = case_params(cfg, np.float64)
    m = ref.PseudoSpeakerVAE(model=dict(input_dim=256, latent_dim=64), classifier=cfg["clf"], optimizer={}, scheduler=dict(T_max=1)).double()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    x, y, eps = case_batch(cfg, 0, np.float64)
    with injected_normals([torch.from_numpy(eps)]):
        loss = m.training_step((torch.from_numpy(x), torch.from_numpy(y)), 0)["loss"]
    loss.backward()
    scal, out, grads = O.train_loss_and_grads(params, x, y, eps)
    assert abs(float(loss) - float(scal["loss"])) < 1e-12
    for k, p in m.named_parameters():
        assert rel_err(grads[k], p.grad.numpy()) < 1e-11, k


def test_torch_port_matches_golden():
    """oracle/torch_port.py (the CPU baseline bench.py times) against the fixtures from the unmodified reference."""
    import torch

    from oracle import torch_port as T

    torch.set_float32_matmul_precision("highest")
    z, cfg = load("train_d256_c2")
    params = case_params(cfg, np.float64)
    mod = T.TorchStep(cfg["D"], cfg["L"], 2).double()
    sd = {}
    for k, v in params.items():
        sd[k.replace("classifier.layers.0", "classifier")] = torch.from_numpy(v)
    mod.load_state_dict(sd)
    opt = torch.optim.Adam(mod.parameters(), lr=1e-3)
    for s in range(cfg.get("steps", 3)):
        x, y, eps = case_batch(cfg, s, np.float64)
        opt.zero_grad()
        loss = mod.loss(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(eps))
        loss.backward()
        assert abs(float(loss) - float(z[f"f64/step{s}/log/train_loss"])) < 1e-12
        opt.step()
        for k, p in mod.named_parameters():
            key = k.replace("classifier.", "classifier.layers.0.") if k.startswith("classifier.") else k
            check_summary(z, f"f64/step{s}/param", key, p.detach().numpy(), 1e-10)


@pytest.mark.parametrize("normalize", [False, True])
def test_vae_backward_for_a_callers_own_loss_vs_torch_autograd(normalize):
    """oracle.vae_backward (gradients for given d loss / d outputs of VAEModel.forward) against torch autograd on the same network
    (nn.Linear / ReLU / exp / F.normalize, the ops of ps_vae/model.py:14-63; the live reference module when /root/reference is present)."""
    import torch

    from oracle.ref_loader import injected_normals, load_reference, reference_available

    D, L, B = 192, 64, 12
    params = O.synth_params(O.vae_param_shapes(D, L), seed=41, dtype=np.float64)
    x, _, eps = O.synth_batch(B, D, L, 2, seed=42, dtype=np.float64)
    rng = np.random.default_rng(43)
    gx, gm, gl = rng.standard_normal((B, D)), rng.standard_normal((B, L)), rng.standard_normal((B, L))
    ours = O.vae_backward(params, x, eps, gx, gm, gl, normalize_decoder=normalize)
    sd = {k[len("model."):]: torch.from_numpy(v) for k, v in params.items()}
    if reference_available():
        model = load_reference().VAEModel(input_dim=D, latent_dim=L, normalize_decoder=normalize).double()
        model.load_state_dict(sd)
        with injected_normals([torch.from_numpy(eps)]):
            x_hat, mu, ls = model(torch.from_numpy(x))
    else:
        def mlp(prefix, h):
            n = len([k for k in sd if k.startswith(prefix) and k.endswith(".weight")])
            for j in range(n):
                h = torch.nn.functional.linear(h, sd[f"{prefix}.{2 * j}.weight"], sd[f"{prefix}.{2 * j}.bias"])
                if j + 1 < n:
                    h = torch.relu(h)
            return h
        for v in sd.values():
            v.requires_grad_(True)
        xt = torch.from_numpy(x)
        mu, ls = mlp("encoder_mu", xt), mlp("encoder_sigma", xt)
        x_hat = mlp("decoder", mu + torch.exp(0.5 * ls) * torch.from_numpy(eps))
        if normalize:
            x_hat = torch.nn.functional.normalize(x_hat, p=2, dim=1)
        model = None
    loss = (x_hat * torch.from_numpy(gx)).sum() + (mu * torch.from_numpy(gm)).sum() + (ls * torch.from_numpy(gl)).sum()
    loss.backward()
    named = dict(model.named_parameters()) if model is not None else sd
    for k, g in ours.items():
        ref = named[k[len("model."):]].grad.numpy()
        assert rel_err(g, ref) <= 1e-12, k
    only_mu = O.vae_backward(params, x, eps, None, gm, None, normalize_decoder=normalize)
    assert all(np.abs(v).max() == 0 for k, v in only_mu.items() if "decoder" in k or "encoder_sigma" in k)     # nothing reaches them


@pytest.mark.parametrize("name", ["transform_prior_threshold", "transform_noise_c3"])
def test_latent_transformation_matches_reference_golden(name):
    """The analysis Langevin variant (analysis/sample_gender_transformation.py:57-99: start from the encoded embedding, PRIOR_WEIGHT,
    per-sample stop at p(y|z) > THRESHOLD) against fixtures produced by running that loop on the unmodified reference modules."""
    from tests.golden_util import case_params, load, rel_err

    z, cfg = load(name)
    for dt, tag, tol in ((np.float64, "f64", 1e-12), (np.float32, "f32", 1e-5)):
        p = case_params(cfg, dt)
        _, mu, _, _ = O.vae_forward(p, z["x"].astype(dt), np.zeros((cfg["N"], cfg["L"]), dt))
        assert rel_err(mu, z[f"{tag}/z_start"]) <= tol
        zf, stop, prob = O.latent_transformation(p, z[f"{tag}/z_start"].astype(dt), list(z["noises"].astype(dt)), cfg["target"], cfg["step_size"],
                                                 cfg["noise_weight"], cfg["prior_weight"], cfg["threshold"])
        assert rel_err(zf, z[f"{tag}/z_final"]) <= tol
        if dt is np.float64:
            assert np.array_equal(stop, z[f"{tag}/stop"])
        assert np.abs(prob - z[f"{tag}/prob"]).max() <= (1e-12 if dt is np.float64 else 1e-5)
    assert len(set(z["f64/stop"].tolist())) > 1


@pytest.mark.parametrize("name", ["embclf_d256_c3", "embclf_d192_c2"])
def test_embedding_classifier_trainer_matches_reference_golden(name):
    """The stand-alone EmbeddingClassifier trainer (embedding_classifier.py:64-100) restated: loss, accuracy, logits and the six gradient
    tensors of step 0 against fixtures from the unmodified module; Adam is pinned separately (adam_cosine.npz)."""
    from tests.golden_util import load, rel_err

    z, cfg = load(name)
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in
         O.synth_params(O.embedding_classifier_param_shapes(cfg["D"], cfg["num_classes"], cfg["hidden_dim"]), seed=cfg["wseed"], dtype=np.float64).items()}
    x, y, _ = O.synth_batch(cfg["B"], cfg["D"], 64, cfg["num_classes"], seed=cfg["dseed"])
    scal, logits, g = O.embedding_classifier_loss_and_grads(p, x.astype(np.float64), y)
    assert abs(float(scal["loss"]) - float(z["f64/step0/log/train_loss"])) <= 1e-12
    assert abs(float(scal["acc"]) - float(z["f64/step0/log/train_acc"])) <= 1e-7
    for k, v in g.items():
        assert rel_err(v, z[f"f64/step0/grad/{k}"]) <= 1e-10, k
    # one Adam step with the module's own optimizer settings reproduces the post-step parameters
    opt = dict(cfg["optimizer"])
    for k in g:
        new, _, _ = O.adam_step(p[k], g[k], np.zeros_like(p[k]), np.zeros_like(p[k]), 1, lr=opt.get("lr", 1e-3), weight_decay=opt.get("weight_decay", 0.0))
        assert rel_err(new, z[f"f64/step0/param/{k}"]) <= 1e-10, k
