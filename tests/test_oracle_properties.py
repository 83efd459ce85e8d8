"""Property tests (hypothesis) of the CPU oracle: size-independent identities of the mixture, boosting-weight and resampling
functions that the GPU parity tests rely on at full batch sizes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import gbnf_oracle as orc

SET = settings(max_examples=60, deadline=None)


def _case(seed, B, C):
    rng = np.random.default_rng(seed)
    logq = (rng.standard_normal((B, C)) * 6 - 20).astype(np.float64)
    rho = rng.uniform(0.05, 1.0, C)
    return rng, logq, rho


@SET
@given(seed=st.integers(0, 10 ** 6), B=st.integers(1, 40), C=st.integers(1, 9), raw=st.booleans(), data=st.data())
def test_flat_mixture_equals_the_recursion(seed, B, C, raw, data):
    """density_experiment.py:612-622 (2-term logsumexp recursion) == logsumexp_c(coef_c + log q_c), incl. the toy skip and the
    raw-rho variant of _rho_gradients (raw rho must stay < 1 for log(1 - rho) to exist)."""
    _, logq, rho = _case(seed, B, C)
    if raw:
        rho = rho * 0.9
    n = data.draw(st.integers(1, C))
    skip = data.draw(st.sampled_from([-1] + list(range(1, n))))
    a = orc.mixture_recursion(logq, rho, n, skip, normalized=not raw)
    b = orc.mixture_flat(logq, rho, n, skip, normalized=not raw)
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)


@SET
@given(seed=st.integers(0, 10 ** 6), B=st.integers(1, 40), C=st.integers(1, 9))
def test_mixture_is_a_convex_combination_and_shifts_with_its_terms(seed, B, C):
    _, logq, rho = _case(seed, B, C)
    G = orc.mixture_recursion(logq, rho, C)
    assert np.all(G <= logq.max(1) + 1e-9) and np.all(G >= logq.min(1) - 1e-9)
    np.testing.assert_allclose(orc.mixture_recursion(logq + 3.25, rho, C), G + 3.25, rtol=1e-10, atol=1e-9)
    # geometric mixture: linear in log q, equal to the common value when all components agree
    Gg = orc.mixture_geometric(logq, rho, C)
    np.testing.assert_allclose(orc.mixture_geometric(2.0 * logq, rho, C), 2.0 * Gg, rtol=1e-10)
    same = np.repeat(logq[:, :1], C, 1)
    np.testing.assert_allclose(orc.mixture_geometric(same, rho, C), logq[:, 0], rtol=1e-10)
    assert np.all(Gg <= G + 1e-9)                   # weighted geometric mean <= weighted arithmetic mean (of the densities)


@SET
@given(seed=st.integers(0, 10 ** 6), B=st.integers(2, 300), mode=st.sampled_from(["density", "toy"]), spread=st.floats(0.01, 30.0))
def test_boosting_weights_are_a_distribution_and_respect_the_clamp(seed, B, mode, spread):
    rng = np.random.default_rng(seed)
    G = (rng.standard_normal(B) * spread - 10).astype(np.float32)
    w = orc.boost_weights(G, mode, batch_size=B)
    assert w.shape == (B,) and np.all(w > 0) and abs(float(w.sum()) - 1.0) < 1e-4
    assert np.argmax(w) == np.argmin(G) or w.max() == w[np.argmin(G)]      # the least likely sample weighs the most
    # invariant to a constant shift of the log-density (softmax)
    np.testing.assert_allclose(orc.boost_weights(G + np.float32(2.5), mode, batch_size=B), w, rtol=2e-4, atol=1e-9)


@SET
@given(seed=st.integers(0, 10 ** 6), B=st.integers(1, 200), n=st.integers(1, 200))
def test_resampling_is_the_left_inverse_cdf(seed, B, n):
    rng = np.random.default_rng(seed)
    w = rng.random(B).astype(np.float32) + 1e-3
    w /= w.sum()
    u = rng.random(n)
    idx = orc.resample_indices(w, u)
    assert idx.shape == (n,) and idx.min() >= 0 and idx.max() < B
    cum = np.cumsum(w.astype(np.float64)); cum /= cum[-1]
    for i, j in zip(u, idx):                        # cum[j-1] < u <= cum[j]
        assert (j == 0 or cum[j - 1] < i) and i <= cum[j] + 1e-15
    order = np.argsort(u)
    assert np.all(np.diff(idx[order]) >= 0)         # monotone in u
