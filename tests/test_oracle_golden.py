"""The CPU oracle against the fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import ALL_CASES as GOLDEN_CASES
from helpers import golden_model, rel_err
from oracle import gbnf_oracle as orc

DENSITY_CASES = [c for c in GOLDEN_CASES if c != "toy_d2"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_component_forward_fp32(golden, name):
    g = golden(name); md = golden_model(g)
    x = g["x"]
    for c in range(md["C"]):
        z, ldj = orc.component_forward(md, x, c)
        assert z.dtype == np.float32
        np.testing.assert_allclose(z, g["z32"][c], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(ldj, g["ldj32"][c], rtol=2e-5, atol=2e-5)
        lq = orc.component_logq(md, x, c)
        assert rel_err(lq, g["logq32"][:, c]) < 5e-6


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_component_forward_fp64_ground_truth(golden, name):
    g = golden(name); md = orc.cast_model(golden_model(g), np.float64)
    lq = orc.all_component_logq(md, g["x"].astype(np.float64))
    assert rel_err(lq, g["logq64"]) < 1e-7   # see log_normal_standard: the reference's .double() run is fp32-contaminated


@pytest.mark.parametrize("name", DENSITY_CASES)
def test_mixture_weights_resample_objective(golden, name):
    g = golden(name); md = golden_model(g)
    x = g["x"]
    tags = sorted({k.rsplit(".", 1)[0] for k in g if k.startswith("kl.")})
    assert tags
    for tag in tags:
        comp = int(tag.split(".")[1][1:]); all_tr = bool(int(tag.split(".")[2][1:]))
        logq = g["logq32"]
        G = orc.mixture_recursion(logq, md["rho"], comp)
        np.testing.assert_allclose(G, g[tag + ".G_ll"], rtol=2e-6, atol=2e-6)
        # flat closed form == recursion
        np.testing.assert_allclose(orc.mixture_flat(logq, md["rho"], comp), g[tag + ".G_ll"], rtol=1e-5, atol=1e-5)
        w = orc.boost_weights(g[tag + ".G_ll"], "density")
        e = np.exp(-g[tag + ".G_ll"] - np.max(-g[tag + ".G_ll"]))
        assert bool(np.max(e / np.sum(e)) > 0.1) == bool(g[tag + ".clamped"])
        np.testing.assert_allclose(w, g[tag + ".w"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(orc.boost_weights(g[tag + ".G_flat"], "density"), g[tag + ".w_flat"], rtol=1e-5, atol=1e-9)
        # resampling: bit-exact on the reference's own weights and uniforms
        assert np.array_equal(orc.resample_indices(g[tag + ".w"], g[tag + ".u"]), g[tag + ".idx"])
        out = orc.compute_kl_pq_loss(md, x, comp, all_tr, u=g[tag + ".u"])
        assert np.array_equal(out["idx"], g[tag + ".idx"])
        for k in ("nll", "G_nll", "g_nll"):
            assert abs(float(out[k]) - float(g[tag + "." + k])) < 2e-5 * max(1.0, abs(float(g[tag + "." + k])))


@pytest.mark.parametrize("name", DENSITY_CASES)
def test_evaluate_and_rho_gradients(golden, name):
    g = golden(name); md = golden_model(g)
    x = g["x"]; B = x.shape[0]
    tags = sorted({k.rsplit(".", 1)[0] for k in g if k.startswith("eval.")})
    for tag in tags:
        comp = int(tag.split(".")[1][1:]); all_tr = bool(int(tag.split(".")[2][1:]))
        ev = orc.evaluate(md, [x[:B // 2], x[B // 2:]], comp, all_tr)
        for k in ("nll", "g_nll", "ratio"):
            assert abs(ev[k] - float(g[tag + "." + k])) < 2e-5 * max(1.0, abs(float(g[tag + "." + k])))
    C = md["C"]
    full = orc.mixture_recursion(g["logq32"], md["rho"], C, normalized=False)
    fixed = orc.mixture_recursion(g["logq32"], md["rho"], C - 1, normalized=False)
    np.testing.assert_allclose(full, g["rhograd.full"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(fixed, g["rhograd.fixed"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g["logq32"][:, C - 1], g["rhograd.new"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.mixture_flat(g["logq32"], md["rho"], C, normalized=False), g["rhograd.full"], rtol=1e-5, atol=1e-5)


def test_toy_objective(golden):
    g = golden("toy_d2"); md = golden_model(g)
    assert md["base_mean"] is not None
    x = g["x"]
    tags = sorted({k.rsplit(".", 1)[0] for k in g if k.startswith("toykl.")})
    assert len(tags) == 2
    for tag in tags:
        comp = int(tag.split(".")[1][1:]); all_tr = bool(int(tag.split(".")[2][1:]))
        bs = int(g[tag + ".batch_size"])
        n, skip = (md["C"], comp) if all_tr else (comp, -1)
        G = orc.mixture_recursion(g["logq32"], md["rho"], n, skip)
        np.testing.assert_allclose(G, g[tag + ".G_ll"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(orc.mixture_flat(g["logq32"], md["rho"], n, skip), g[tag + ".G_ll"], rtol=1e-5, atol=1e-5)
        w = orc.boost_weights(g[tag + ".G_ll"], "toy", bs)
        np.testing.assert_allclose(w, g[tag + ".w"], rtol=1e-5, atol=1e-9)
        assert np.array_equal(orc.resample_indices(g[tag + ".w"], g[tag + ".u"]), g[tag + ".idx"])
        out = orc.compute_kl_pq_loss(md, x, comp, all_tr, u=g[tag + ".u"], mode="toy", batch_size=bs)
        for k in ("nll", "G_nll", "g_nll"):
            assert abs(float(out[k]) - float(g[tag + "." + k])) < 2e-5 * max(1.0, abs(float(g[tag + "." + k])))


def test_grid_density_geometric_mixture(golden):
    """plot_boosted_fwd_flow_density (utils/density_plotting.py:185-226) run by the reference on a 12 x 12 grid."""
    g = golden("toy_d2"); md = golden_model(g)
    zz, C = g["grid.zz"], md["C"]
    logq = orc.all_component_logq(md, zz)
    for c in range(C):
        np.testing.assert_allclose(np.exp(logq[:, c]).reshape(12, 12), g[f"grid.prob.c{c}"], rtol=2e-5, atol=1e-9)
    total = np.exp(orc.mixture_geometric(logq, md["rho"], C)).reshape(12, 12)
    np.testing.assert_allclose(total, g["grid.total_prob"], rtol=2e-5, atol=1e-9)
    # a component with rho == 0 is left out of the sum but not of the normaliser (:200-201, :222)
    rho = md["rho"].copy(); rho[2] = 0.0
    ref = sum(logq[:, c] * rho[c] for c in range(C) if c != 2) / rho.sum()
    np.testing.assert_allclose(orc.mixture_geometric(logq, rho, C), ref, rtol=1e-6)


def test_properties():
    rng = np.random.default_rng(0)
    logq = rng.standard_normal((500, 6)).astype(np.float32) * 5 - 30
    rho = np.array([1, .5, .25, .125, .0625, .05], dtype=np.float32)
    coef = orc.mixture_log_coefficients(rho, 6)
    assert abs(np.sum(np.exp(coef)) - 1.0) < 1e-12                      # simplex weights
    np.testing.assert_allclose(np.exp(coef), rho.astype(np.float64) / rho.astype(np.float64).sum(), rtol=1e-12)
    w = orc.boost_weights(orc.mixture_recursion(logq, rho, 6))
    assert abs(w.sum() - 1) < 1e-5 and w.min() > 0
    u = np.sort(rng.random(500))
    idx = orc.resample_indices(w, u)
    assert np.all(np.diff(idx) >= 0) and idx.min() >= 0 and idx.max() < 500
    assert orc.sample_component(rho, 3, 0.0) == 0 and orc.sample_component(rho, 3, 0.999999) == 2
    assert orc.sample_component(rho, 6, 0.3, exclude=0) != 0
    # synthetic-model constructor gives a usable model
    md = orc.make_synthetic_model("glow", 7, 2, 2, 16, seed=3, init_rows=64)
    lq = orc.all_component_logq(md, rng.standard_normal((5, 7)).astype(np.float32))
    assert lq.shape == (5, 2) and np.all(np.isfinite(lq))
    md2 = orc.unflatten_model(orc.flatten_model(md))
    np.testing.assert_array_equal(orc.all_component_logq(md2, np.ones((3, 7), np.float32)),
                                  orc.all_component_logq(md, np.ones((3, 7), np.float32)))


def test_torch_timing_variant_matches_numpy_oracle(golden):
    import torch
    for name in ("glow_d43", "realnvp_d6_bn", "realnvp_d5_mixed", "glow_d6_additive_relu"):
        g = golden(name); md = golden_model(g)
        tm = orc.to_torch_model(md)
        x = torch.from_numpy(g["x"])
        with torch.no_grad():
            lq = torch.stack([orc.torch_component_logq(tm, x, c) for c in range(md["C"])], 1).numpy()
            G, w = orc.torch_density_step(tm, x)
        assert rel_err(lq, g["logq32"]) < 5e-6
        Gn = orc.mixture_recursion(g["logq32"], md["rho"], md["C"])
        np.testing.assert_allclose(G.numpy(), Gn, rtol=5e-6, atol=5e-6)
        np.testing.assert_allclose(w.numpy(), orc.boost_weights(Gn), rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_inverse_round_trip_and_reference_decode(golden, name):
    """The oracle's inverse direction: exact inverse of the reference's forward (fp64), and equal to the reference's own decode
    where upstream's decode runs (1-D Glow, additive coupling: FlowStep.decode models/glow.py:344-366)."""
    g = golden(name); md = golden_model(g)
    m64 = orc.cast_model(md, np.float64)
    for c in range(md["C"]):
        x, ldj = orc.component_inverse(m64, g["z64"][c], c)
        np.testing.assert_allclose(x, g["x"].astype(np.float64), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(ldj, -g["ldj64"][c], rtol=1e-6, atol=1e-6)
        if f"dec64.c{c}" in g:
            np.testing.assert_allclose(x, g[f"dec64.c{c}"], rtol=1e-12, atol=1e-12)
            x32, _ = orc.component_inverse(md, g["z32"][c], c)
            np.testing.assert_allclose(x32, g[f"dec32.c{c}"], rtol=2e-5, atol=2e-5)
    assert any(k.startswith("dec64.") for k in golden("glow_d6_additive_relu"))


def test_oracle_component_assignment_matches_sample_component():
    rng = np.random.default_rng(3)
    rho = rng.random(7).astype(np.float32) + 0.05
    u = rng.random(500)
    got = orc.assign_components(rho, 7, u)
    assert all(got[i] == orc.sample_component(rho, 7, float(u[i])) for i in range(500))
    got = orc.assign_components(rho, 7, u, exclude=3)
    assert 3 not in got and all(got[i] == orc.sample_component(rho, 7, float(u[i]), exclude=3) for i in range(500))


@pytest.mark.parametrize("name", ["glow_d43", "realnvp_d6_bn"])
def test_oracle_update_rho_loop_vs_reference(golden, name):
    """The reference's own update_rho loop (run by make_golden.py with the two undefined log-message names of
    models/boosted_flow.py:185 defined in the fixture script only)."""
    g = golden(name); md = golden_model(g)
    B = g["x"].shape[0]
    rho = orc.update_rho(md, [g["x"][:B // 2], g["x"][B // 2:]], md["C"] - 1, int(g["urho.iters"]), float(g["urho.lr"]))
    np.testing.assert_allclose(rho, g["urho.rho_final"], rtol=2e-5, atol=2e-6)
    assert not np.allclose(rho, md["rho"])
