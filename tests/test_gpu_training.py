"""Training step of the NEW component through the library (-m gpu), SURVEY 8 a12 / 8(f).1: the forward of the component under
training runs the fixed-component kernel, loss.backward() runs gbnf_component_backward (recompute-in-kernel); gradients must
match the reference's autograd -- here: torch autograd through the host mirror's reference-equal forward (pinned to the
reference's outputs by tests/test_host_modules.py) in fp64 and fp32 -- to 1e-4."""
import argparse
import copy

import numpy as np
import pytest
import torch

import gbnf_b200
from helpers import build_model, golden_model
from oracle import gbnf_oracle as orc

pytestmark = pytest.mark.gpu

CASES = {
    "glow_affine_d43_h512": dict(kind="glow", D=43, C=3, K=5, h=512),
    "glow_affine_d43_h64": dict(kind="glow", D=43, C=2, K=2, h=64),
    "glow_additive_relu_d6_h128": dict(kind="glow", D=6, C=2, K=3, h=128, coupling="additive", act="relu"),
    "glow_affine_relu_d21_h256": dict(kind="glow", D=21, C=2, K=4, h=256, act="relu"),
    "glow_affine_d64_h384": dict(kind="glow", D=64, C=2, K=2, h=384),
    "glow_affine_d2_h128": dict(kind="glow", D=2, C=2, K=2, h=128),
    # RealNVP without BatchNorm: two networks per step, halves swapped on the flipped steps (odd D: |z1| alternates)
    "realnvp_tanh_d6_h256": dict(kind="realnvp", D=6, C=3, K=5, h=256),
    "realnvp_mixed_d5_h64": dict(kind="realnvp", D=5, C=2, K=4, h=64, act="mixed"),
    "realnvp_relu_d2_h256": dict(kind="realnvp", D=2, C=2, K=1, h=256, act="relu"),
    "realnvp_tanh_d43_h512": dict(kind="realnvp", D=43, C=2, K=3, h=512),
}


def _autograd_reference(model, c, x, dtype):
    """Gradients by torch autograd through forward_autograd (== the reference's forward) in `dtype`."""
    flow = copy.deepcopy(model.flows[c]).to(dtype)
    if model.component_type == "glow":
        for s_src, s_dst in zip(model.flows[c].steps(), flow.steps()):
            s_dst.permutation.set_indices(s_src.permutation.indices)
            s_dst.actnorm.inited = True
    for p in flow.parameters():
        p.requires_grad_(True)
    z, ldj = flow.forward_autograd(x.to(dtype))
    g_nll = -((-0.5 * np.log(2 * np.pi) - 0.5 * z.pow(2)).sum(1) + ldj)
    g_nll.mean().backward()
    return [p.grad.detach().double() for p in flow.parameters()], g_nll.detach().double()


@pytest.mark.parametrize("mode", ["fp32", "f16fast"])
@pytest.mark.parametrize("name", list(CASES))
def test_new_component_gradients_match_autograd(name, mode):
    kw = dict(CASES[name])
    md = orc.make_synthetic_model(kw.pop("kind"), kw.pop("D"), kw.pop("C"), kw.pop("K"), kw.pop("h"), seed=13, **kw)
    B = 300
    x = torch.from_numpy(np.random.default_rng(4).standard_normal((B, md["D"])).astype(np.float32)).cuda()
    model = build_model(md, "cuda", gemm_mode=mode)
    try:
        c = md["C"] - 1
        model.component = c
        for i, f in enumerate(model.flows):
            for p in f.parameters():
                p.requires_grad_(i == c)                       # init_boosted_lr: only the new component trains
        model.train()
        ref64, nll64 = _autograd_reference(model, c, x, torch.float64)
        z, _, _, ldj, _ = model(x=x, components="c")
        assert z.grad_fn is not None and type(z.grad_fn).__name__.startswith("_FusedComponentFlow")
        g_nll = -((-0.5 * np.log(2 * np.pi) - 0.5 * z.pow(2)).sum(1) + ldj)
        loss = g_nll.mean()
        loss.backward()
        model.check_status()
        tol_v = 1e-5 if mode == "fp32" else 1e-4
        assert abs(float(loss.detach()) - float(nll64.mean())) <= tol_v * abs(float(nll64.mean())) + (0 if md["D"] >= 16 else 2e-2 * (mode != "fp32"))
        for p, r in zip(model.flows[c].parameters(), ref64):
            assert p.grad is not None and p.grad.shape == r.shape
            scale = float(r.abs().max()) + 1e-12
            err = float((p.grad.double() - r).abs().max())
            # the backward sweep is fp32 in every mode; in the f16 modes the upstream gradient dz = z / B comes from the
            # tensor-core forward's z (fp16 operands, ~1e-3), which bounds the gradient's accuracy there
            gtol = 1e-4 if mode == "fp32" else 5e-3
            assert err <= gtol * scale + 1e-7, (name, tuple(r.shape), err, scale)
        for i, f in enumerate(model.flows):
            if i != c:
                assert all(p.grad is None for p in f.parameters())
    finally:
        model.release()


def test_training_step_through_the_driver_function_and_optimizer():
    """compute_kl_pq_loss(...)['nll'].backward() + optimizer.step() (density_experiment.py:359-374): the loss goes down and the
    re-tiled component follows its parameters."""
    md = orc.make_synthetic_model("glow", 10, 3, 3, 128, seed=2)
    model = build_model(md, "cuda", gemm_mode="f16fast")
    try:
        model.component, model.all_trained = 2, False
        for i, f in enumerate(model.flows):
            for p in f.parameters():
                p.requires_grad_(i == 2)
        model.train()
        opt = torch.optim.Adam([p for p in model.flows[2].parameters()], lr=2e-3)
        x = torch.randn(512, 10, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) * 0.5 + 0.3
        args = argparse.Namespace(flow="boosted")
        gen = torch.Generator(device="cuda").manual_seed(11)
        first = last = None
        for it in range(30):
            opt.zero_grad()
            losses = gbnf_b200.compute_kl_pq_loss(model, x, args, generator=gen)
            losses["nll"].backward()
            torch.nn.utils.clip_grad_norm_(model.flows[2].parameters(), 5.0)
            opt.step()
            first = float(losses["g_nll"]) if first is None else first
            last = float(losses["g_nll"])
        assert np.isfinite(last) and last < first - 0.05, (first, last)
        model.check_status()
    finally:
        model.release()


def test_unsupported_configurations_fall_back_to_autograd():
    """RealNVP WITH BatchNorm: train-mode batch statistics couple the rows of a batch -> the caller's autograd."""
    md = orc.make_synthetic_model("realnvp", 6, 2, 3, 128, seed=3, batch_norm=True)
    model = build_model(md, "cuda", gemm_mode="fp32")
    try:
        for p in model.flows[1].parameters():
            p.requires_grad_(True)
        model.component = 1
        x = torch.randn(64, 6, device="cuda")
        z, _, _, ldj, _ = model(x=x, components="c")
        (z.pow(2).sum() + ldj.sum()).backward()
        assert all(p.grad is not None for p in model.flows[1].parameters())
        assert not type(z.grad_fn).__name__.startswith("_FusedComponentFlow")
    finally:
        model.release()
