"""Host-side model of `tc_act4` (csrc/coupling_tc.cuh): the accurate tanh of four values with one reciprocal.
fp32 arithmetic restated in numpy (exact exp2 / reciprocal stand in for ex2.approx / rcp.approx, both <= 2 ulp); pins the error bound
DESIGN.md §4.1 quotes, the cap (no overflow of the four-fold product) and the special values.  The kernel itself is checked on the
GPU against the fp64 oracle (tests/test_gpu_parity.py, mode "f16")."""
import numpy as np

F = np.float32
C2 = F(2.8853900817779268)        # 2 log2(e)
CAP = F(28.853900817779268)       # 20 log2(e): 2 v capped at 20


def tanh4(v):
    v = np.asarray(v, dtype=F).reshape(-1, 4)
    with np.errstate(over="ignore", invalid="ignore"):
        t = np.where(np.isnan(v), v, np.minimum(v * C2, CAP)).astype(F)        # min.NaN: a NaN passes the cap
        a = (np.exp2(t).astype(F) + F(1)).astype(F)
        p01, p23 = a[:, 0] * a[:, 1], a[:, 2] * a[:, 3]
        rm = F(-2) * (F(1) / (p01 * p23)).astype(F)
        r01, r23 = rm * p23, rm * p01
        out = np.stack([r01 * a[:, 1] + F(1), r01 * a[:, 0] + F(1), r23 * a[:, 3] + F(1), r23 * a[:, 2] + F(1)], 1)
    return out.astype(F), (p01 * p23)


def test_error_bound_and_no_overflow():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.standard_normal(400000) * 3, rng.uniform(-30, 30, 400000), rng.uniform(-1e-2, 1e-2, 40000)]).astype(F)
    out, prod = tanh4(v)
    ref = np.tanh(v.reshape(-1, 4).astype(np.float64))
    assert np.abs(out - ref).max() < 5e-7
    assert np.isfinite(prod).all() and prod.max() < 6e34
    # what the kernel stores is the fp16 rounding: identical to the rounding of the exact value except on rounding ties
    h, hr = out.astype(np.float16), ref.astype(np.float16)
    assert (h != hr).mean() < 5e-3 and np.abs(h.astype(F) - hr.astype(F)).max() <= 2 ** -11 + 1e-7


def test_special_values():
    out, _ = tanh4([np.inf, -np.inf, 0.0, 1e4, -1e4, 88.0, 88.0, 88.0])
    np.testing.assert_allclose(out.ravel(), [1, -1, 0, 1, -1, 1, 1, 1], atol=2e-7)
    out, _ = tanh4([np.nan, 0.5, 0.5, 0.5])
    assert np.isnan(out[0, 0])          # NaN is not swallowed by the cap (it may spread to the other three values of the same row)
