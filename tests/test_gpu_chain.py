"""GPU tests of the chained decoder kernel (csrc/decoder_chain.cuh; option "decode_chain"): its own file, run in its own process -- a
tcgen05 / mbarrier protocol bug traps the CUDA context, which must not take the rest of the suite with it."""
import numpy as np
import pytest
import torch

from oracle import philox_ref as PR
from oracle import ps_vae_oracle as O
from tests.golden_util import case_params, rel_err

pytestmark = pytest.mark.gpu

DECODE_CHAIN_DEFAULT = 1       # on by default since round 2 (validated by this file, +28 % on unconditional sampling)


def _gu():
    from tests import gpu_util

    return gpu_util


def _module(G, D=256, H=None, seed=3):
    cfg = dict(D=D, L=64, wseed=seed, clf=dict(input_dim=64, num_classes=2))
    if H is not None:
        cfg.update(H=H, nh=2)
        shapes = O.vae_param_shapes(D, 64, H, 2) + O.classifier_param_shapes(64, 2)
        params = {k: v.astype(np.float32) for k, v in O.synth_params(shapes, seed=seed, dtype=np.float64).items()}
        return G.module_from_cfg(cfg, "bf16", params=params), cfg, params
    return G.module_from_cfg(cfg, "bf16"), cfg, case_params(cfg, np.float32)


@pytest.mark.parametrize("N", [1, 300, 256, 257, 2 * 74 * 256 + 13, 5 * 74 * 256])
def test_chained_decoder_matches_layer_by_layer_path(N):
    """Unconditional decode (z from Philox): the chained kernel against the per-layer GEMM path -- same z, same bf16 rounding points."""
    G = _gu()
    module, cfg, _ = _module(G)
    hot = module.hot_path
    try:
        G.L.set_option("decode_chain", 0)
        hot.manual_seed(11, 0)
        ref, zr = hot.decode(None, num_samples=N, return_z=True, row0=5)
        G.L.set_option("decode_chain", 1)
        hot.manual_seed(11, 0)
        got, zg = hot.decode(None, num_samples=N, return_z=True, row0=5)
        torch.cuda.synchronize()
    finally:
        G.L.set_option("decode_chain", DECODE_CHAIN_DEFAULT)
    assert torch.equal(zr, zg)                                   # same Philox counters: bit-identical z
    e = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
    assert torch.isfinite(got).all() and e <= 1e-5, e


@pytest.mark.parametrize("D,H", [(256, 512), (192, 256), (64, 128), (256, 384)])
def test_chained_decoder_vs_oracle_shapes(D, H):
    """Given z (the conditional-sampling tail) at several widths, against the numpy oracle on bf16-rounded operands."""
    G = _gu()
    from oracle import bf16_twin as T

    module, cfg, params = _module(G, D, H, seed=5)
    N = 777
    z = PR.philox_normal(N, 64, 3, 1, 0).astype(np.float32)
    zt = torch.from_numpy(z).to(G.DEV)
    try:
        G.L.set_option("decode_chain", 1)
        got = module.hot_path.decode(zt)
        torch.cuda.synchronize()
    finally:
        G.L.set_option("decode_chain", DECODE_CHAIN_DEFAULT)
    want, _ = T._mlp_fwd({k: v.astype(np.float64) for k, v in params.items()}, "model.decoder", T.bf16_round(z))
    # a hidden activation whose fp32-accumulated value sits on a bf16 rounding boundary lands one bf16 ulp from the fp64-accumulating twin's
    assert rel_err(got.cpu().numpy(), want) <= 2e-4


def test_chained_decoder_shards_are_bit_identical():
    """Row i of the output does not depend on how the rows are sharded (Philox counter = global row): 3 shards == one batch."""
    G = _gu()
    import pseudo_speaker_vae_b200 as P

    module, cfg, _ = _module(G)
    N = 1000
    try:
        G.L.set_option("decode_chain", 1)
        module.hot_path.manual_seed(77, 0)
        full = P.sample_on_device(module, N)
        parts = []
        for r in range(3):
            row0, rows = P.shard_rows(N, r, 3)
            module.hot_path.manual_seed(77, 0)
            parts.append(P.sample_on_device(module, rows, row0=row0))
        torch.cuda.synchronize()
    finally:
        G.L.set_option("decode_chain", DECODE_CHAIN_DEFAULT)
    assert torch.equal(torch.cat(parts), full)


@pytest.mark.parametrize("B", [300, 256 * 74 * 2 + 77, 65536])
@pytest.mark.parametrize("xdt", ["fp32", "bf16"])
def test_train_chain_matches_layer_by_layer_step(B, xdt):
    """Option train_chain: the decoder half of the training forward pass as one chained kernel (decoder_chain.cuh, TRAIN) against the per-layer
    GEMM path: same losses (fp32 summation order), same gradients, and -- at the small batch -- the bf16-emulating twin."""
    G = _gu()
    from oracle import bf16_twin as T

    cfg = dict(D=256, L=64, wseed=3, clf=dict(input_dim=64, num_classes=2))
    module = G.module_from_cfg(cfg, "bf16")
    hot = module.hot_path
    x, y, eps = O.synth_batch(B, 256, 64, 2, seed=31)
    xt = torch.from_numpy(x).to(G.DEV)
    if xdt == "bf16":
        xt = xt.to(torch.bfloat16)
    yt, et = torch.from_numpy(y).to(G.DEV), torch.from_numpy(eps).to(G.DEV)

    def run():
        g = torch.empty(hot.arena.numel, device=G.DEV)
        losses, _, _ = hot.step(xt, yt, et, grads=g)
        torch.cuda.synchronize()
        return g, losses

    try:
        G.L.set_option("train_chain", 0)
        g0, l0 = run()
        G.L.set_option("train_chain", 1)
        n0 = G.L.lib().psvae_launch_count()
        g1, l1 = run()
        n1 = G.L.lib().psvae_launch_count()
    finally:
        G.L.set_option("train_chain", 0)
    assert torch.allclose(l0, l1, rtol=2e-6, atol=1e-7), (l0[:4], l1[:4])
    e = ((g1.double() - g0.double()).norm() / g0.double().norm()).item()
    assert e <= 1e-5, e
    gd0, gd1 = G.flat_to_dict(module, g0), G.flat_to_dict(module, g1)
    worst = max(rel_err(gd1[k], gd0[k]) for k in gd0)
    assert worst <= 2e-5, {k: f"{rel_err(gd1[k], gd0[k]):.1e}" for k in gd0}
    if B == 300:
        xin = xt.float().cpu().numpy()
        scal, out, grads = T.train_loss_and_grads_bf16(case_params(cfg, np.float32), xin, y, eps)
        assert abs(float(l1[0]) - float(scal["loss"])) <= 2e-5
        assert max(rel_err(gd1[k], grads[k]) for k in grads) <= 1e-3
    print(f"train_chain B={B} x={xdt}: launches {n1 - n0}, grad diff {e:.1e}, worst tensor {worst:.1e}")
