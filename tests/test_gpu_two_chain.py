"""GPU parity tests of the two-chain schedule (csrc/coupling_tc4.cuh): Glow / affine / tanh components of hidden width 512.
It must be BITWISE identical to the single-chain pipelined kernel (same operands, same accumulation order) and inside the
f16 gates against the fp64 oracle.  GBNF_TC4 is read when the handle is created: 0 = single chain, 2 = prefer two chains."""
import os

import numpy as np
import pytest
import torch

from helpers import build_model, golden_model
from oracle import gbnf_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _model(md, mode, tc4):
    old = os.environ.get("GBNF_TC4")
    try:
        if tc4 is None:
            os.environ.pop("GBNF_TC4", None)
        else:
            os.environ["GBNF_TC4"] = str(tc4)
        m = build_model(md, "cuda", gemm_mode=mode)
        m.pack_all()          # creates the handle (the switch is read here)
        return m
    finally:
        if old is None:
            os.environ.pop("GBNF_TC4", None)
        else:
            os.environ["GBNF_TC4"] = old


def _run(md, x, mode, tc4, n_mix=None):
    m = _model(md, mode, tc4)
    try:
        G, lq = m.mixture_log_density(x, md["C"] if n_mix is None else n_mix, return_logq=True)
        m.check_status()
        torch.cuda.synchronize()
        return G.clone(), lq.clone(), m.info()
    finally:
        m.release()


@pytest.mark.parametrize("mode", ["f16fast", "f16"])
def test_bench_shape_bitwise_equal_to_single_chain(mode):
    """cfg3's shape at a ragged 65 536 + 77 rows: 6.9 work units per CTA, chains in different units at unit boundaries."""
    md = orc.make_synthetic_model("glow", 43, 8, 5, 512, seed=1)
    x = torch.from_numpy(np.random.default_rng(5).standard_normal((65536 + 77, 43)).astype(np.float32)).cuda()
    G2, lq2, inf2 = _run(md, x, mode, None)
    G1, lq1, inf1 = _run(md, x, mode, 0)
    assert inf2["two_chain"] == 1 and inf1["two_chain"] == 0 and inf2["pipelined"] == 1
    assert torch.equal(lq2, lq1) and torch.equal(G2, G1)
    ref = orc.all_component_logq(orc.cast_model(md, np.float64), x[:1024].cpu().numpy().astype(np.float64))
    err = np.abs(lq2[:1024].cpu().numpy() - ref) / np.abs(ref)
    assert err.max() < 1e-4, err.max()


@pytest.mark.parametrize("C,K,D,B", [(2, 1, 12, 300), (4, 2, 43, 777), (6, 3, 21, 1000), (8, 2, 32, 129), (4, 1, 5, 1), (4, 2, 63, 200)])
def test_small_shapes_forced_two_chain(C, K, D, B):
    """Few tiles (the cost model would split down to one component per unit): GBNF_TC4=2 forces the two-chain kernel.
    D = 63: two z tiles do not fit next to a three-stage weight ring -> the handle keeps the single-chain kernel."""
    md = orc.make_synthetic_model("glow", D, C, K, 512, seed=10 + C)
    x = torch.from_numpy(np.random.default_rng(C).standard_normal((B, D)).astype(np.float32)).cuda()
    G2, lq2, inf2 = _run(md, x, "f16fast", 2)
    G1, lq1, inf1 = _run(md, x, "f16fast", 0)
    assert inf2["two_chain"] == (1 if D <= 44 else 0) and inf1["two_chain"] == 0
    assert torch.equal(lq2, lq1) and torch.equal(G2, G1)
    ref = orc.all_component_logq(orc.cast_model(md, np.float64), x.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(lq2.cpu().numpy(), ref, rtol=1e-4, atol=2e-2 if D < 16 else 2e-3)
    Gref = orc.mixture_recursion(ref, md["rho"].astype(np.float64), C)
    np.testing.assert_allclose(G2.cpu().numpy(), Gref, rtol=1e-4, atol=2e-2 if D < 16 else 2e-3)


def test_partial_mixture_and_odd_counts_fall_back():
    """n_mix < C with an even count stays on two chains; an odd count of components runs the single-chain kernel."""
    md = orc.make_synthetic_model("glow", 43, 8, 2, 512, seed=3)
    x = torch.from_numpy(np.random.default_rng(9).standard_normal((4096, 43)).astype(np.float32)).cuda()
    m = _model(md, "f16fast", 2)
    m0 = _model(md, "f16fast", 0)
    try:
        for n_mix, two in ((6, 1), (7, 0), (2, 1), (1, 0)):
            G = m.mixture_log_density(x, n_mix)
            assert m.info()["two_chain"] == two, (n_mix, m.info())
            assert torch.equal(G, m0.mixture_log_density(x, n_mix))
        m.check_status()
    finally:
        m.release(); m0.release()


def test_reference_golden_h512():
    """The reference-made fixture glow_d43_h512 (C = 2, K = 2) through the two-chain kernel."""
    g = dict(np.load(os.path.join(GOLDEN, "glow_d43_h512.npz")))
    md = golden_model(g)
    x = torch.from_numpy(g["x"]).cuda()
    G, lq, inf = _run(md, x, "f16fast", 2)
    assert inf["two_chain"] == 1
    ref = g["logq64"]
    assert np.max(np.abs(lq.cpu().numpy() - ref) / np.abs(ref)) < 1e-4
