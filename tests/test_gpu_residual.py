"""GPU parity of ResidualNet coupling networks (models/layers.py:246-301, RealNVP with args.coupling_network == 'residual')
against fixtures made by the unmodified reference: exact fp32 kernel only; the f16 tensor-core modes must refuse loudly."""
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
    assert md["act"] == "residual"
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


@pytest.mark.parametrize("mode", ["f16", "f16fast"])
def test_residual_rejected_by_tensor_core_modes(golden, mode):
    md = golden_model(golden("realnvp_d6_residual"))
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
