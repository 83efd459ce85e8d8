"""Failure-mode tests (-m gpu): non-finite terms through the logsumexp kernels follow torch.logsumexp
(density_experiment.py:571,622), values that fp16 cannot hold are REPORTED instead of silently turning into a density of 1,
the toy base density is copied at gbnf_set_base time, and the pack cache notices every way a parameter can change."""
import numpy as np
import pytest
import torch

import gbnf_b200
from gbnf_b200._lib import GbnfError
from helpers import build_model
from oracle import gbnf_oracle as orc

pytestmark = pytest.mark.gpu


def _model(mode, **kw):
    cfg = dict(kind="glow", D=8, C=3, K=2, h=128)
    cfg.update(kw)
    md = orc.make_synthetic_model(cfg.pop("kind"), cfg.pop("D"), cfg.pop("C"), cfg.pop("K"), cfg.pop("h"), seed=5, **cfg)
    return build_model(md, "cuda", gemm_mode=mode), md


@pytest.mark.parametrize("C", [3, 8, 20])          # 128-bit register path (<= 16 terms) and the online path
def test_mixture_kernel_follows_torch_logsumexp_on_non_finite_terms(C):
    model, md = _model("fp32", C=C)
    try:
        rng = np.random.default_rng(C)
        lq = rng.standard_normal((64, C)).astype(np.float32) * 3 - 20
        lq[3, 1] = np.nan
        lq[5, :] = -np.inf
        lq[7, 0] = -np.inf
        lq[9, C - 1] = np.inf
        lq[11, :] = np.nan
        G = model.mixture_from_logq(torch.from_numpy(lq).cuda(), C).cpu()
        coef = torch.from_numpy(orc.mixture_log_coefficients(md["rho"].astype(np.float64), C)).float()
        ref = torch.logsumexp(torch.from_numpy(lq) + coef, dim=1)                     # the reference's ATen op
        assert torch.isnan(G[3]) and torch.isnan(G[11]) and torch.isnan(ref[3])
        assert G[5] == -np.inf and ref[5] == -np.inf
        assert G[9] == np.inf and ref[9] == np.inf
        ok = torch.isfinite(ref)
        np.testing.assert_allclose(G[ok].numpy(), ref[ok].numpy(), rtol=2e-6, atol=2e-6)
        assert ok[7] and ok.sum() == 64 - 4
    finally:
        model.release()


@pytest.mark.parametrize("mode", ["fp32", "f16", "f16fast"])
def test_nan_input_row_gives_nan_density_not_density_one(mode):
    model, md = _model(mode)
    try:
        x = torch.randn(300, md["D"], device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        x[17, :] = float("nan")
        G, lq = model.mixture_log_density(x, md["C"], return_logq=True)
        assert torch.isnan(G[17]) and torch.isnan(lq[17]).all(), (G[17], lq[17])
        ok = torch.ones(300, dtype=torch.bool, device="cuda"); ok[17] = False
        assert torch.isfinite(G[ok]).all()
        if mode == "fp32":
            model.check_status()
        else:                                   # the fp16 operand conversion saw a non-finite value: reported, once
            with pytest.raises(GbnfError, match="not finite in fp16") as ei:
                model.check_status()
            assert ei.value.code == -4
            model.check_status()
    finally:
        model.release()


@pytest.mark.parametrize("kind", ["glow", "realnvp"])
def test_fp16_range_overflow_is_reported(kind):
    """Un-standardised data (|x| > 65504) cannot enter an fp16 GEMM operand: GBNF_ERR_NUMERIC, surfaced by the NEXT call (no
    synchronisation on the hot path) or at once by check_status(); the fp32 mode evaluates the same rows."""
    model, md = _model("f16fast", kind=kind, D=6)
    ref, _ = _model("fp32", kind=kind, D=6)
    try:
        x = torch.randn(256, 6, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
        model.mixture_log_density(x, md["C"]); model.check_status()
        x[100] = 1.0e6
        model.mixture_log_density(x, md["C"])
        torch.cuda.synchronize()
        with pytest.raises(GbnfError, match="not finite in fp16"):
            model.mixture_log_density(x[:10].contiguous(), md["C"])        # first call after the offending launch finished
        model.mixture_log_density(x[:10].contiguous(), md["C"]); model.check_status()      # cleared
        assert torch.isfinite(ref.mixture_log_density(x, md["C"])[:100]).all(); ref.check_status()
    finally:
        model.release(); ref.release()


def test_relu_activation_overflow_is_reported():
    model, md = _model("f16", kind="glow", D=6, act="relu", h=128)
    try:
        with torch.no_grad():
            model.flows[0].flow.layers[0].block.network[0].bias.fill_(7.0e4)      # hidden pre-activations ~ 7e4 > fp16 max
        x = torch.randn(200, 6, device="cuda")
        model.mixture_log_density(x, 1)
        with pytest.raises(GbnfError, match="not finite in fp16"):
            model.check_status()
    finally:
        model.release()


def test_toy_base_is_copied_and_tracked():
    md = orc.make_synthetic_model("realnvp", 2, 2, 1, 128, seed=3, rho_init="uniform", toy_base=True)
    model = build_model(md, "cuda", gemm_mode="fp32")
    try:
        x = torch.randn(128, 2, device="cuda")
        a = model.component_log_density(x).clone()
        with torch.no_grad():
            model.base_dist_mean.add_(0.5)                 # in-place edit of the buffer the handle was created from
        b = model.component_log_density(x)
        md2 = gbnf_b200.extract_model(model, toy_base=True)
        np.testing.assert_allclose(b.cpu().numpy(), orc.all_component_logq(md2, x.cpu().numpy()), rtol=1e-5, atol=1e-5)
        assert not torch.allclose(a, b)
    finally:
        model.release()


def test_invalidate_picks_up_data_writes():
    model, md = _model("fp32")
    try:
        x = torch.randn(64, md["D"], device="cuda")
        a = model.component_log_density(x, 0, 1).clone()
        model.flows[0].flow.layers[0].block.network[0].weight.data.mul_(1.25)      # .data write: no version bump
        model.invalidate()
        b = model.component_log_density(x, 0, 1)
        assert not torch.equal(a, b)
        np.testing.assert_allclose(b.cpu().numpy()[:, 0], orc.component_logq(gbnf_b200.extract_model(model), x.cpu().numpy(), 0),
                                   rtol=1e-5, atol=1e-5)
    finally:
        model.release()


def test_entry_points_restore_current_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    md = orc.make_synthetic_model("glow", 6, 2, 2, 64, seed=1)
    model = build_model(md, "cuda:1", gemm_mode="fp32")
    try:
        torch.cuda.set_device(0)
        x = torch.randn(64, 6, device="cuda:1")
        model.component_log_density(x)
        assert torch.cuda.current_device() == 0
        assert torch.empty(1, device="cuda").device.index == 0
    finally:
        model.release()
