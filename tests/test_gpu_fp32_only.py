"""GPU parity of the two model options the exact fp32 kernel alone serves -- ResidualNet coupling networks (models/layers.py:246-301,
RealNVP with args.coupling_network == 'residual') and Glow's invertible 1x1 convolution as the permutation (models/glow.py:275-278,
models/layers.py:722-796) -- against fixtures made by the unmodified reference; the f16 tensor-core modes must refuse them loudly."""
import numpy as np
import pytest
import torch

import gbnf_b200
from conftest import FP32_ONLY_CASES
from helpers import build_model, golden_model
from oracle import gbnf_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", FP32_ONLY_CASES)
def test_residual_logq_mixture_weights_vs_reference(golden, name):
    g = golden(name); md = golden_model(g)
    assert md["act"] == "residual" or md["components"][0]["steps"][0]["ic"] is not None
    model = build_model(md, "cuda", gemm_mode="fp32")
    try:
        x = torch.from_numpy(g["x"]).cuda()
        lq = model.component_log_density(x).cpu().numpy()
        np.testing.assert_allclose(lq, g["logq64"], rtol=1e-5, atol=2e-5)
        for c in range(md["C"]):
            with torch.no_grad():
                z, _, _, ldj, _ = model(x=x, components=c)
            np.testing.assert_allclose(z.cpu().numpy(), g["z64"][c], rtol=2e-5, atol=2e-5)
            np.testing.assert_allclose(ldj.cpu().numpy(), g["ldj64"][c], rtol=2e-5, atol=2e-5)
        # fused mixture + boosting weights against the reference's own G_ll / w
        tags = sorted({k.rsplit(".", 1)[0] for k in g if k.startswith("kl.")})
        assert tags
        for tag in tags:
            comp = int(tag.split(".")[1][1:])
            if comp == 0:
                continue
            G = model.mixture_log_density(x, comp)
            np.testing.assert_allclose(G.cpu().numpy(), g[tag + ".G_ll"], rtol=2e-5, atol=2e-5)
            w = model.boosting_weights(G)
            np.testing.assert_allclose(w.cpu().numpy(), g[tag + ".w"], rtol=2e-4, atol=1e-8)
        # inverse direction: decode(encode(x)) == x
        z, ldj = model.component_forward(x, 0)
        xr, ldji = model.component_inverse(z, 0)
        assert float((xr - x).abs().max()) < 2e-4 and float((ldj + ldji).abs().max()) < 2e-3
        model.check_status()
    finally:
        model.release()


@pytest.mark.parametrize("name", ["realnvp_d6_residual", "glow_d6_invconv"])
@pytest.mark.parametrize("mode", ["f16", "f16fast"])
def test_rejected_by_tensor_core_modes(golden, mode, name):
    md = golden_model(golden(name))
    model = build_model(md, "cuda", gemm_mode=mode)
    with pytest.raises(gbnf_b200.GbnfError):
        model.component_log_density(torch.zeros(4, md["D"], device="cuda"))


def test_residual_synthetic_vs_oracle_large_batch():
    """h = 256, two blocks, 20 000 rows (several CTAs per SM wave, ragged tail) against the fp64 oracle."""
    md = orc.make_synthetic_model("realnvp", 6, 3, 4, 256, seed=5, depth=2, act="residual")
    x = np.random.default_rng(3).standard_normal((20000 + 13, 6)).astype(np.float32)
    model = build_model(md, "cuda", gemm_mode="fp32")
    try:
        lq = model.component_log_density(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = orc.all_component_logq(orc.cast_model(md, np.float64), x.astype(np.float64))
        np.testing.assert_allclose(lq, ref, rtol=1e-5, atol=5e-5)
    finally:
        model.release()


def test_invconv_synthetic_vs_oracle_large_batch():
    """D = 43, h = 512 (32-row tiles), plain weight matrices, 10 000 + 7 rows against the fp64 oracle; decode(encode(x)) == x."""
    md = orc.make_synthetic_model("glow", 43, 2, 3, 512, seed=8, invconv=True)
    x = np.random.default_rng(4).standard_normal((10007, 43)).astype(np.float32)
    model = build_model(md, "cuda", gemm_mode="fp32")
    try:
        xd = torch.from_numpy(x).cuda()
        lq = model.component_log_density(xd).cpu().numpy()
        ref = orc.all_component_logq(orc.cast_model(md, np.float64), x.astype(np.float64))
        np.testing.assert_allclose(lq, ref, rtol=1e-5, atol=5e-5)
        z, ldj = model.component_forward(xd, 1)
        xr, ldji = model.component_inverse(z, 1)
        assert float((xr - xd).abs().max()) < 5e-4 and float((ldj + ldji).abs().max()) < 2e-3
    finally:
        model.release()
