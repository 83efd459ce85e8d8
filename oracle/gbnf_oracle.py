"""CPU oracle for the GBNF boosted-mixture density path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm
(robert-giaquinto/gradient-boosted-normalizing-flows) for the hot path named in
BASELINE.json: per-component log q_c(x), rho-weighted logsumexp mixture, boosting
weights, inverse-CDF resampling and the sample-weighted objective.

It is the *checker*.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product package
(gradient-boosted-normalizing-flows_b200/) never does.

Parity status: PINNED.  tests/golden/*.npz were produced by running the unmodified
reference (imported from /root/reference by tests/golden/make_golden.py) and
tests/test_oracle_golden.py checks every function below against them.

Every function works in the dtype of its inputs (float32 = reference numerics,
float64 = ground truth).  All file:line citations are relative to the reference tree.

Model container ("model dict"), produced by the package's extract_model():
    {'kind': 'glow'|'realnvp', 'D', 'h', 'K', 'C', 'depth',
     'act': 'tanh'|'relu'|'mixed', 'coupling': 'affine'|'additive',
     'rho': [C], 'base_mean': [D] | None, 'base_scale': [D] | None,
     'components': [{'flip_init': int, 'steps': [step, ...]}, ...]}
    glow step    : {'an_bias': [D], 'an_logs': [D], 'perm': int64 [D], 'net': [(W, b), ...]}
    realnvp step : {'bn': None | {'log_gamma','beta','mean','var'}, 't': [(W, b), ...], 's': [(W, b), ...]}
    W follows nn.Linear: [out_features, in_features].
"""
import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)
BN_EPS = 1e-5  # models/layers.py:323


# --------------------------------------------------------------------------------------
# A.1  coupling MLPs                                   models/layers.py:208-243
# --------------------------------------------------------------------------------------
def _act(a, kind):
    if kind == "tanh":
        return np.tanh(a)
    if kind == "relu":
        return np.maximum(a, 0)
    raise ValueError(kind)


def mlp(x, layers, act):
    """Linear -> [act, Linear] * depth -> act -> Linear   (TanhNet :230-243 / ReLUNet :208-227).
    act == 'residual': ResidualNet (models/layers.py:246-301) -- layers = [initial, (block first, block second) * n, final];
    t = initial(x); t = t + second(relu(first(relu(t)))) per block; final(t)."""
    if act == "residual":
        t = x @ layers[0][0].T + layers[0][1]
        for i in range(1, len(layers) - 1, 2):
            (Wa, ba), (Wb, bb) = layers[i], layers[i + 1]
            u = np.maximum(t, 0) @ Wa.T + ba
            t = t + (np.maximum(u, 0) @ Wb.T + bb)
        return t @ layers[-1][0].T + layers[-1][1]
    a = x @ layers[0][0].T + layers[0][1]
    for W, b in layers[1:]:
        a = _act(a, act) @ W.T + b
    return a


# --------------------------------------------------------------------------------------
# A.2  Glow 1-D step                                   models/glow.py:317-342
# --------------------------------------------------------------------------------------
def invconv_weight(ic):
    """InvertibleConv1x1.get_weight(reverse=False) for a feature vector (h * w == 1): (W [D, D], dlogdet).
    models/layers.py:751-779.  ic = {'weight'} (plain) or {'p', 'lower', 'upper', 'log_s', 'sign_s'} (LU-decomposed)."""
    if "weight" in ic:
        W = ic["weight"]
        return W, np.linalg.slogdet(W.astype(np.float64))[1].astype(W.dtype)
    D = ic["log_s"].shape[0]
    dt = ic["lower"].dtype
    l_mask = np.tril(np.ones((D, D), dtype=dt), -1)
    lower = ic["lower"] * l_mask + np.eye(D, dtype=dt)
    u = ic["upper"] * l_mask.T + np.diag(ic["sign_s"] * np.exp(ic["log_s"]))
    return ic["p"] @ (lower @ u), np.sum(ic["log_s"])


def glow_step(z, ldj, step, coupling="affine", act="tanh"):
    D = z.shape[1]
    h0 = D // 2
    # 1. ActNorm1d: center then scale, logdet += sum(logs)       models/layers.py:488-518
    y = (z + step["an_bias"]) * np.exp(step["an_logs"])
    ldj = ldj + np.sum(step["an_logs"])
    if step.get("ic") is not None:
        # 2'. InvertibleConv1x1 (F.conv2d with a [D, D, 1, 1] kernel on a 1 x 1 image): z_i = sum_j W_ij y_j   layers.py:781-796
        W, dlogdet = invconv_weight(step["ic"])
        y = y @ W.T
        ldj = ldj + dlogdet
    else:
        # 2. Permute1d: output column j = input column indices[j]     models/layers.py:661-668
        y = y[:, step["perm"]]
    # 3. coupling                                                 models/glow.py:326-340
    z1, z2 = y[:, :h0], y[:, h0:]
    hh = mlp(z1, step["net"], act)
    if coupling == "additive":
        z2 = z2 + hh
    else:
        shift, raw = hh[:, 0::2], hh[:, 1::2]                    # "cross" split utils/utilities.py:155-156
        scale = 1.0 / (1.0 + np.exp(-(raw + 2.0)))               # torch.sigmoid(scale + 2.)
        z2 = (z2 + shift) * scale
        ldj = ldj + np.sum(np.log(scale), axis=1)                # torch.log(scale) literal, glow.py:338
    return np.concatenate([z1, z2], axis=1), ldj


# --------------------------------------------------------------------------------------
# A.3  RealNVP step                                    models/transformations.py:560-579
# --------------------------------------------------------------------------------------
def batchnorm_eval(x, bn):
    """Eval-mode RealNVP BatchNorm (running statistics).  models/layers.py:349-358."""
    x_hat = (x - bn["mean"]) / np.sqrt(bn["var"] + x.dtype.type(BN_EPS))
    y = np.exp(bn["log_gamma"]) * x_hat + bn["beta"]
    ladj = bn["log_gamma"] - 0.5 * np.log(bn["var"] + x.dtype.type(BN_EPS))
    return y, np.sum(ladj)


def realnvp_step(x, ldj, step, flipped, act="tanh"):
    D = x.shape[1]
    h0 = D // 2
    if step.get("bn") is not None:
        x, bn_ldj = batchnorm_eval(x, step["bn"])
    else:
        bn_ldj = 0.0
    if flipped:                                                   # transformations.py:568-571
        z2, z1 = x[:, :h0], x[:, h0:]
    else:
        z1, z2 = x[:, :h0], x[:, h0:]
    t_act = "relu" if act == "mixed" else act                     # models/realnvp.py:47-51
    s_act = "tanh" if act == "mixed" else act
    shift = mlp(z1, step["t"], t_act)
    scale = mlp(z1, step["s"], s_act)
    z2 = shift + z2 * np.exp(scale)
    z = np.concatenate([z1, z2], axis=1)                          # [z1, z2'] for BOTH flips (:576)
    return z, ldj + np.sum(scale, axis=1) + bn_ldj


# --------------------------------------------------------------------------------------
# component forward: Glow.encode glow.py:92-110, RealNVPFlow.encode realnvp.py:115-127
# --------------------------------------------------------------------------------------
def component_forward(model, x, c):
    comp = model["components"][c]
    z = x
    ldj = np.zeros(x.shape[0], dtype=x.dtype)
    for k, step in enumerate(comp["steps"]):
        if model["kind"] == "glow":
            z, ldj = glow_step(z, ldj, step, model.get("coupling", "affine"), model.get("act", "tanh"))
        else:
            flipped = ((k + comp["flip_init"]) % 2) > 0           # realnvp.py:38,119
            z, ldj = realnvp_step(z, ldj, step, flipped, model.get("act", "tanh"))
    return z, ldj


# --------------------------------------------------------------------------------------
# inverse direction (sampling): x = f_c^{-1}(z).  Upstream: Glow.decode glow.py:112-123 -> FlowNet.decode :254-260 ->
# FlowStep.decode :344-366 (its affine branch sums over dim=[1,2,3] and raises on 1-D data; the additive branch runs and
# is pinned by a golden fixture); RealNVPFlow.decode realnvp.py:97-113 / transformations.py:581-599 pairs the wrong halves
#  and is NOT an inverse of encode.  The functions below are the exact mathematical inverses of the forward
# steps above, in the reference's own order of operations (coupling^-1, permutation^-1, ActNorm^-1 / BatchNorm^-1).
# --------------------------------------------------------------------------------------
def glow_step_inverse(y, ldj, step, coupling="affine", act="tanh"):
    D = y.shape[1]
    h0 = D // 2
    z1, z2 = y[:, :h0], y[:, h0:]
    hh = mlp(z1, step["net"], act)
    if coupling == "additive":
        z2 = z2 - hh                                              # glow.py:350
    else:
        shift, raw = hh[:, 0::2], hh[:, 1::2]
        scale = 1.0 / (1.0 + np.exp(-(raw + 2.0)))
        z2 = z2 / scale - shift                                   # glow.py:354-356
        ldj = ldj - np.sum(np.log(scale), axis=1)                 # glow.py:357
    z = np.concatenate([z1, z2], axis=1)
    if step.get("ic") is not None:
        W, dlogdet = invconv_weight(step["ic"])
        zin = z @ np.linalg.inv(W.astype(np.float64)).astype(z.dtype).T      # layers.py:772-776, 791-795
        ldj = ldj - dlogdet
    else:
        zin = np.empty_like(z)
        zin[:, step["perm"]] = z                                  # inverse of y = z[:, indices], layers.py:661-668
    x = zin * np.exp(-step["an_logs"]) - step["an_bias"]          # ActNorm reverse: scale then center, layers.py:505-518
    return x, ldj - np.sum(step["an_logs"])


def realnvp_step_inverse(z, ldj, step, flipped, act="tanh"):
    D = z.shape[1]
    h0 = D // 2
    n1 = D - h0 if flipped else h0                                # forward output is [z1, z2'] for both flips
    z1, z2 = z[:, :n1], z[:, n1:]
    t_act = "relu" if act == "mixed" else act
    s_act = "tanh" if act == "mixed" else act
    shift = mlp(z1, step["t"], t_act)
    scale = mlp(z1, step["s"], s_act)
    z2 = (z2 - shift) * np.exp(-scale)
    x = np.concatenate([z2, z1], axis=1) if flipped else np.concatenate([z1, z2], axis=1)
    ldj = ldj - np.sum(scale, axis=1)
    bn = step.get("bn")
    if bn is not None:                                            # inverse of batchnorm_eval
        x_hat = (x - bn["beta"]) * np.exp(-bn["log_gamma"])
        x = x_hat * np.sqrt(bn["var"] + x.dtype.type(BN_EPS)) + bn["mean"]
        ldj = ldj - np.sum(bn["log_gamma"] - 0.5 * np.log(bn["var"] + x.dtype.type(BN_EPS)))
    return x, ldj


def component_inverse(model, z, c):
    """x = f_c^{-1}(z) and the log-det of the INVERSE map (= -log_det_j of the forward map at x)."""
    comp = model["components"][c]
    x = z
    ldj = np.zeros(z.shape[0], dtype=z.dtype)
    for k in range(len(comp["steps"]) - 1, -1, -1):
        step = comp["steps"][k]
        if model["kind"] == "glow":
            x, ldj = glow_step_inverse(x, ldj, step, model.get("coupling", "affine"), model.get("act", "tanh"))
        else:
            flipped = ((k + comp["flip_init"]) % 2) > 0
            x, ldj = realnvp_step_inverse(x, ldj, step, flipped, model.get("act", "tanh"))
    return x, ldj


def assign_components(rho, n, u, exclude=-1):
    """Per-sample component ids for mixture sampling under the inverse-CDF rule of sample_component:
    j_i = #{k : cum_k < u_i}, cum = cumsum_fp64(rho[:n]) / sum."""
    p = np.asarray(rho[:n], dtype=np.float64).copy()
    if exclude >= 0:
        p[exclude] = 0.0
    cum = np.cumsum(p) / np.sum(p)
    return np.minimum(np.searchsorted(cum, np.asarray(u, dtype=np.float64), side="left"), n - 1).astype(np.int64)


# --------------------------------------------------------------------------------------
# A.4  base densities
# --------------------------------------------------------------------------------------
def log_normal_standard(z):
    """sum_D(-0.5 log 2pi - 0.5 z^2)                    utils/distributions.py:44-60.
    Note: the reference evaluates -0.5*log(2*PI) on a float32 tensor (:47-52) and starts log-det at float32 zeros
    (glow.py:93), so even a .double() copy of the reference carries ~1e-8 relative float32 contamination; this
    function is exact in the dtype of z, which is why the fp64 golden values are compared at 1e-7, not 1e-12."""
    c = z.dtype.type(-0.5) * np.log(z.dtype.type(2.0) * z.dtype.type(math.pi))
    return np.sum(c - z.dtype.type(0.5) * z * z, axis=-1)


def log_normal_diag_base(z, mean, scale):
    """torch.distributions.Normal(mean, scale).log_prob(z).sum(1)  (toy path: generative_flow.py:38-42,
    toy_experiment.py:424).  Normal.log_prob = -((z-m)^2)/(2 s^2) - log s - log sqrt(2 pi)."""
    var = scale * scale
    lp = -((z - mean) ** 2) / (2 * var) - np.log(scale) - z.dtype.type(math.log(math.sqrt(2 * math.pi)))
    return np.sum(lp, axis=1)


def component_logq(model, x, c):
    z, ldj = component_forward(model, x, c)
    if model.get("base_mean") is not None:
        return log_normal_diag_base(z, model["base_mean"].astype(z.dtype), model["base_scale"].astype(z.dtype)) + ldj
    return log_normal_standard(z) + ldj


def all_component_logq(model, x, n=None):
    n = model["C"] if n is None else n
    if n == 0:
        return np.zeros((x.shape[0], 0), dtype=x.dtype)
    return np.stack([component_logq(model, x, c) for c in range(n)], axis=1)


# --------------------------------------------------------------------------------------
# A.5  mixture recursion
# --------------------------------------------------------------------------------------
def _lse2(a, b):
    m = np.maximum(a, b)
    return m + np.log(np.exp(a - m) + np.exp(b - m))


def mixture_recursion(logq, rho, n, skip_c=-1, normalized=True):
    """2-term logsumexp recursion over components 0..n-1.

    density path: density_experiment.py:612-622 (train, n=component) / :561-571 (eval, n=component+1)
    toy path    : toy_experiment.py:413-432 with `continue` on c == skip_c (all_trained second pass); the
                  skipped component's rho stays inside the running normaliser rho[0:c+1].
    normalized=False reproduces BoostedFlow._rho_gradients (models/boosted_flow.py:124-134) which uses raw rho[c].
    n == 0 returns zeros (density_experiment.py:612)."""
    B = logq.shape[0]
    dt = logq.dtype
    rho = np.asarray(rho, dtype=dt)
    G = np.zeros(B, dtype=dt)
    for c in range(n):
        if c == skip_c:
            continue
        if c == 0:
            G = logq[:, 0].copy()
        else:
            r = rho[c] / np.sum(rho[: c + 1]) if normalized else rho[c]
            G = _lse2(np.log(dt.type(1.0) - r) + G, np.log(r) + logq[:, c])
    return G


def mixture_log_coefficients(rho, n, skip_c=-1, normalized=True):
    """Flat form of the recursion: G = logsumexp_c(coef[c] + logq[:, c]) with
    coef[c] = log r_c + sum_{j>c, j != skip} log(1 - r_j), r_0 := 1.  Components that do not take part get -inf.
    Computed in float64; this is the quantity the CUDA mixture kernel consumes."""
    rho = np.asarray(rho, dtype=np.float64)
    coef = np.full(len(rho), -np.inf)
    r = np.ones(len(rho))
    for c in range(1, n):
        r[c] = rho[c] / np.sum(rho[: c + 1]) if normalized else rho[c]
    for c in range(n):
        if c == skip_c:
            continue
        v = math.log(r[c])
        for j in range(c + 1, n):
            if j != skip_c:
                v += math.log(1.0 - r[j])
        coef[c] = v
    return coef


def mixture_flat(logq, rho, n, skip_c=-1, normalized=True):
    coef = mixture_log_coefficients(rho, n, skip_c, normalized).astype(logq.dtype)
    act = np.isfinite(coef)
    if not act.any():
        return np.zeros(logq.shape[0], dtype=logq.dtype)
    t = logq[:, : len(coef)][:, act] + coef[act]
    m = np.max(t, axis=1)
    return m + np.log(np.sum(np.exp(t - m[:, None]), axis=1))


def mixture_geometric(logq, rho, n):
    """Log of the 'all components' grid density of plot_boosted_fwd_flow_density (utils/density_plotting.py:199-226):
    total = sum_c logq[:, c] * rho[c] over the first n components with rho[c] != 0 (accumulated in component order, in the
    dtype of logq), divided by sum(rho[0:n]).  The reference plots exp() of this."""
    dt = logq.dtype
    rho = np.asarray(rho, dtype=dt)
    total = np.zeros(logq.shape[0], dtype=dt)
    for c in range(n):
        if rho[c] == 0.0:
            continue
        total = total + logq[:, c] * rho[c]
    return total / np.sum(rho[:n])


# --------------------------------------------------------------------------------------
# A.6  boosting weights
# --------------------------------------------------------------------------------------
def boost_weights(G_ll, mode="density", batch_size=None):
    """density: density_experiment.py:627-641 (+ utils/utilities.py:12-14); toy: toy_experiment.py:440,453-459."""
    dt = G_ll.dtype
    u = -G_ll
    e = np.exp(u - np.max(u))
    w = e / np.sum(e)
    if mode == "density":
        if np.max(w) > dt.type(0.1):
            w = np.maximum(np.minimum(w, dt.type(0.1)), dt.type(0.01))
        s = np.sum(w)
        if s != dt.type(1.0):
            w = w / s
    else:
        w = w / np.sum(w)
        if np.max(w) > dt.type(0.1):
            w = np.maximum(np.minimum(w, dt.type(0.1)), dt.type(0.1 / batch_size))
            w = w / np.sum(w)
    return w


# --------------------------------------------------------------------------------------
# A.7  resampling / assignment contract: idx = #{k : cum_k < u}, cum = cumsum_fp64(w)/sum_fp64(w)
# equals torch.multinomial(w, n>=2, replacement=True) replayed with torch.rand(n, dtype=float64)
# (density_experiment.py:643-644).
# --------------------------------------------------------------------------------------
def resample_indices(w, u):
    cum = np.cumsum(np.asarray(w, dtype=np.float64))
    cum = cum / cum[-1]
    idx = np.searchsorted(cum, np.asarray(u, dtype=np.float64), side="left")
    return np.minimum(idx, len(cum) - 1).astype(np.int64)


def sample_component(rho, n, u, exclude=-1):
    """Component id under the same inverse-CDF rule for '1:c' / '1:c-1' / '-c' (models/boosted_flow.py:76-91)."""
    p = np.asarray(rho, dtype=np.float64).copy()
    if exclude >= 0:
        p[exclude] = 0.0
        n = len(p)
    return int(resample_indices(p[:n], np.array([u]))[0])


# --------------------------------------------------------------------------------------
# A.8  objective / evaluation                          density_experiment.py:606-674, :544-603
# --------------------------------------------------------------------------------------
def compute_kl_pq_loss(model, x, component, all_trained, u=None, mode="density", batch_size=None):
    """Returns dict(nll, G_nll, g_nll, weights, idx).  `u`: float64 uniforms, one per row, for the resampling."""
    C = model["C"]
    cnew = min(component, C - 1)                                  # _sample_component("c") boosted_flow.py:72-74
    boosted_branch = (all_trained or component > 0) if mode == "density" else component > 0
    if not boosted_branch:
        g = -component_logq(model, x, cnew)
        return {"nll": np.mean(g), "g_nll": np.mean(g), "G_nll": x.dtype.type(0.0), "weights": None, "idx": None}
    if mode == "density":
        n, skip = component, -1
    else:
        n, skip = (C if all_trained else component), (component if all_trained else -1)
    logq = all_component_logq(model, x, n)
    G_ll = mixture_recursion(logq, model["rho"], n, skip)
    w = boost_weights(G_ll, mode, batch_size if batch_size is not None else x.shape[0])
    idx = resample_indices(w, u)
    g = -component_logq(model, x[idx], cnew)
    return {"nll": np.mean(g), "G_nll": np.mean(-G_ll), "g_nll": np.mean(g), "weights": w, "idx": idx, "G_ll": G_ll}


def evaluate(model, batches, component, all_trained):
    """density_experiment.py:544-603: full mixture over component+1 comps + the 'c' component alone."""
    C = model["C"]
    G_nll, g_nll = [], []
    for x in batches:
        n = component + 1
        logq = all_component_logq(model, x, n)
        G_nll.append(-mixture_recursion(logq, model["rho"], n))
        if component > 0 or all_trained:
            g_nll.append(-component_logq(model, x, min(component, C - 1)))
    G_nll = np.concatenate(G_nll)
    out = {"nll": float(np.mean(G_nll))}
    if component > 0 or all_trained:
        g_nll = np.concatenate(g_nll)
        out["g_nll"] = float(np.mean(g_nll))
        out["ratio"] = float(np.mean(g_nll - G_nll))
    else:
        out["g_nll"] = out["nll"]
        out["ratio"] = 0.0
    return out


# --------------------------------------------------------------------------------------
# ActNorm data-dependent init (models/layers.py:473-486) -- used to build synthetic models
# --------------------------------------------------------------------------------------
def update_rho(model, batches, component, rho_iters, rho_lr):
    """Decayed-step SGD on rho[component]: models/boosted_flow.py:141-207 (approximate=True branch).  Per mini-batch:
    new / fixed / full log-likelihoods with the RAW-rho recursion of _rho_gradients (:119-139), gradient =
    mean((-g_ll) - (-G_ll)) with g_ll = new, G_ll = fixed (:183-184), step = lr / (0.05 i + 1), rho clamped to [0.01, 100]
    (:193-194), stop when i > 10 and (i > iters or |delta| < 0.001) (:203-204).  Returns the updated rho vector."""
    rho = np.array(model["rho"], dtype=np.float32).copy()
    prev = float(rho[component])
    n = component + 1
    for i in range(rho_iters):
        x = batches[i % len(batches)]
        m = dict(model); m["rho"] = rho
        lq = all_component_logq(m, x, n)
        fixed = mixture_recursion(lq, rho, n - 1, normalized=False) if n > 1 else np.zeros(x.shape[0], dtype=x.dtype)
        new = lq[:, n - 1] if n > 1 else np.zeros(x.shape[0], dtype=x.dtype)
        gradient = float(np.mean((-new) - (-fixed)))
        step = rho_lr / (0.05 * i + 1)
        r = min(max(prev - step * gradient, 0.01), 100.0)
        rho[component] = r
        dif, prev = abs(prev - r), r
        if i > 10 and (i > rho_iters or dif < 0.001):
            break
    return rho


def actnorm_init(sample, scale=1.0):
    bias = -np.mean(sample, axis=0)
    var = np.mean((sample + bias) ** 2, axis=0)
    logs = np.log(scale / (np.sqrt(var) + sample.dtype.type(1e-6)))
    return bias.astype(sample.dtype), logs.astype(sample.dtype)


# --------------------------------------------------------------------------------------
# model-dict helpers (casting, npz round trip, synthetic construction)
# --------------------------------------------------------------------------------------
def cast_model(model, dtype):
    def cv(v):
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            return v.astype(dtype)
        if isinstance(v, dict):
            return {k: cv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return type(v)(cv(x) for x in v)
        return v
    return cv(model)


def flatten_model(model):
    """model dict -> flat {name: ndarray} for np.savez."""
    flat = {"meta.kind": np.array(model["kind"]), "meta.act": np.array(model.get("act", "tanh")),
            "meta.coupling": np.array(model.get("coupling", "affine")),
            "meta.dims": np.array([model["D"], model["h"], model["K"], model["C"], model.get("depth", 1)], dtype=np.int64),
            "rho": np.asarray(model["rho"])}
    if model.get("base_mean") is not None:
        flat["base_mean"], flat["base_scale"] = model["base_mean"], model["base_scale"]
    for c, comp in enumerate(model["components"]):
        flat[f"c{c}.flip_init"] = np.array(comp["flip_init"], dtype=np.int64)
        for k, st in enumerate(comp["steps"]):
            p = f"c{c}.k{k}."
            if model["kind"] == "glow":
                flat[p + "an_bias"], flat[p + "an_logs"] = st["an_bias"], st["an_logs"]
                if st.get("ic") is not None:
                    for n_, v in st["ic"].items():
                        flat[p + "ic." + n_] = v
                else:
                    flat[p + "perm"] = st["perm"]
                nets = {"net": st["net"]}
            else:
                if st.get("bn") is not None:
                    for n_, v in st["bn"].items():
                        flat[p + "bn." + n_] = v
                nets = {"t": st["t"], "s": st["s"]}
            for nn_, layers in nets.items():
                for i, (W, b) in enumerate(layers):
                    flat[p + f"{nn_}.W{i}"], flat[p + f"{nn_}.b{i}"] = W, b
    return flat


def unflatten_model(flat):
    D, h, K, C, depth = [int(v) for v in flat["meta.dims"]]
    model = {"kind": str(flat["meta.kind"]), "act": str(flat["meta.act"]), "coupling": str(flat["meta.coupling"]),
             "D": D, "h": h, "K": K, "C": C, "depth": depth, "rho": flat["rho"],
             "base_mean": flat["base_mean"] if "base_mean" in flat else None,
             "base_scale": flat["base_scale"] if "base_scale" in flat else None, "components": []}
    for c in range(C):
        steps = []
        for k in range(K):
            p = f"c{c}.k{k}."
            def layers(nn_):
                n = 0                      # however many Linear layers were stored (depth + 2; ResidualNet: 2 depth + 2)
                while p + f"{nn_}.W{n}" in flat:
                    n += 1
                return [(flat[p + f"{nn_}.W{i}"], flat[p + f"{nn_}.b{i}"]) for i in range(n)]
            if model["kind"] == "glow":
                ic = {n_[len(p) + 3:]: v for n_, v in flat.items() if n_.startswith(p + "ic.")}
                steps.append({"an_bias": flat[p + "an_bias"], "an_logs": flat[p + "an_logs"],
                              "perm": flat[p + "perm"] if not ic else None, "ic": ic or None, "net": layers("net")})
            else:
                bn = None
                if p + "bn.mean" in flat:
                    bn = {n_: flat[p + "bn." + n_] for n_ in ("log_gamma", "beta", "mean", "var")}
                steps.append({"bn": bn, "t": layers("t"), "s": layers("s")})
        model["components"].append({"flip_init": int(flat[f"c{c}.flip_init"]), "steps": steps})
    return model


def _linear_init(rng, out_f, in_f):
    """nn.Linear default init: U(+-1/sqrt(fan_in)) for weight (kaiming_uniform a=sqrt(5)) and bias."""
    bound = 1.0 / math.sqrt(in_f)
    W = rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    b = rng.uniform(-bound, bound, size=(out_f,)).astype(np.float32)
    return W, b


def make_synthetic_model(kind, D, C, K, h, seed=1, depth=1, act="tanh", coupling="affine", batch_norm=False,
                         rho_init="decreasing", toy_base=False, init_rows=4096, x_init=None, invconv=False):
    """Random-init model of the named architecture with the same init *distributions* as the reference
    (nn.Linear default init; ActNorm initialised from data; permutation = shuffled reversed arange;
    rho per models/boosted_flow.py:32-39).  Used for full-size parity and the benchmark; NOT bit-identical to a
    torch-seeded reference model (tests/golden holds those)."""
    rng = np.random.default_rng(seed)
    if rho_init == "decreasing":
        rho = np.maximum(1.0 / np.power(2.0, np.arange(C, dtype=np.float32)), 0.05).astype(np.float32)
    else:
        rho = np.full(C, 1.0 / C, dtype=np.float32)
    model = {"kind": kind, "D": D, "h": h, "K": K, "C": C, "depth": depth, "act": act, "coupling": coupling,
             "rho": rho, "base_mean": None, "base_scale": None, "components": []}
    if toy_base:
        model["base_mean"] = (0.1 * rng.standard_normal(D)).astype(np.float32)
        model["base_scale"] = np.full(D, 3.0, dtype=np.float32)
    h0 = D // 2
    h1 = D - h0
    if x_init is None:
        x_init = rng.standard_normal((init_rows, D)).astype(np.float32)
    for c in range(C):
        steps = []
        z = x_init
        ldj = np.zeros(z.shape[0], dtype=np.float32)
        for k in range(K):
            if kind == "glow":
                out = 2 * h1 if coupling == "affine" else h1
                dims = [h0] + [h] * (depth + 1) + [out]
                net = [_linear_init(rng, dims[i + 1], dims[i]) for i in range(depth + 2)]
                perm = np.arange(D - 1, -1, -1)[rng.permutation(D)].astype(np.int64)
                bias, logs = actnorm_init(z)
                st = {"an_bias": bias, "an_logs": logs, "perm": perm, "net": net}
                if invconv:     # plain (not LU-decomposed) weight: a random rotation (upstream's init) times mild column scales
                    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
                    st["ic"] = {"weight": (q * np.exp(0.1 * rng.standard_normal(D))).astype(np.float32)}
                    st["perm"] = None
                z, ldj = glow_step(z, ldj, st, coupling, act)
            else:
                flipped = ((k + c) % 2) > 0
                i_d, o_d = (h1, h0) if flipped else (h0, h1)
                nlin = 2 * depth + 2 if act == "residual" else depth + 2      # ResidualNet: initial, 2 per block, final
                dims = [i_d] + [h] * (nlin - 1) + [o_d]
                t = [_linear_init(rng, dims[i + 1], dims[i]) for i in range(nlin)]
                s = [_linear_init(rng, dims[i + 1], dims[i]) for i in range(nlin)]
                if act == "residual":      # a block's second Linear starts at uniform(-1e-3, 1e-3) (models/layers.py:262-264)
                    for net_ in (t, s):
                        for i in range(2, nlin - 1, 2):
                            net_[i] = (rng.uniform(-1e-3, 1e-3, net_[i][0].shape).astype(np.float32),
                                       rng.uniform(-1e-3, 1e-3, net_[i][1].shape).astype(np.float32))
                bn = None
                if batch_norm and k < K - 1:
                    bn = {"log_gamma": (0.1 * rng.standard_normal(D)).astype(np.float32),
                          "beta": (0.1 * rng.standard_normal(D)).astype(np.float32),
                          "mean": z.mean(0).astype(np.float32), "var": z.var(0).astype(np.float32)}
                st = {"bn": bn, "t": t, "s": s}
                z, ldj = realnvp_step(z, ldj, st, flipped, act)
            steps.append(st)
        model["components"].append({"flip_init": c, "steps": steps})
    return model


# --------------------------------------------------------------------------------------
# torch-CPU execution of the same restatement (timing arm only).
# The reference IS PyTorch ops on the CPU (nn.Linear -> addmm/MKL, torch.tanh, torch.logsumexp); numpy's tanh and
# BLAS are slower than torch's, so bench.py's reference arm / cpu_baseline time THIS variant to give the reference
# the speed it really has on the host.  tests/test_oracle_golden.py checks it against the numpy functions above.
# --------------------------------------------------------------------------------------
def to_torch_model(model):
    import torch

    def cv(v):
        if isinstance(v, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(v))
        if isinstance(v, dict):
            return {k: cv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return type(v)(cv(x) for x in v)
        return v
    return cv(model)


def torch_component_logq(tm, x, c):
    """Same math as component_logq(), torch CPU tensors, no autograd (call under torch.no_grad())."""
    import torch
    import torch.nn.functional as F
    comp = tm["components"][c]
    D = x.shape[1]
    h0 = D // 2
    act = tm.get("act", "tanh")

    def net(v, layers, a):
        fn = torch.tanh if a == "tanh" else torch.relu
        v = F.linear(v, layers[0][0], layers[0][1])
        for W, b in layers[1:]:
            v = F.linear(fn(v), W, b)
        return v
    z = x
    ldj = torch.zeros(x.shape[0], dtype=x.dtype)
    for k, st in enumerate(comp["steps"]):
        if tm["kind"] == "glow":
            y = (z + st["an_bias"]) * torch.exp(st["an_logs"])
            ldj = ldj + st["an_logs"].sum()
            y = y[:, st["perm"]]
            z1, z2 = y[:, :h0], y[:, h0:]
            hh = net(z1, st["net"], act)
            if tm.get("coupling", "affine") == "additive":
                z2 = z2 + hh
            else:
                scale = torch.sigmoid(hh[:, 1::2] + 2.0)
                z2 = (z2 + hh[:, 0::2]) * scale
                ldj = ldj + torch.log(scale).sum(1)
            z = torch.cat([z1, z2], 1)
        else:
            if st.get("bn") is not None:
                bn = st["bn"]
                z = torch.exp(bn["log_gamma"]) * ((z - bn["mean"]) / torch.sqrt(bn["var"] + BN_EPS)) + bn["beta"]
                ldj = ldj + (bn["log_gamma"] - 0.5 * torch.log(bn["var"] + BN_EPS)).sum()
            if ((k + comp["flip_init"]) % 2) > 0:
                z2, z1 = z[:, :h0], z[:, h0:]
            else:
                z1, z2 = z[:, :h0], z[:, h0:]
            shift = net(z1, st["t"], "relu" if act == "mixed" else act)
            scale = net(z1, st["s"], "tanh" if act == "mixed" else act)
            z = torch.cat([z1, shift + z2 * torch.exp(scale)], 1)
            ldj = ldj + scale.sum(1)
    if tm.get("base_mean") is not None:
        var = tm["base_scale"] ** 2
        lp = -((z - tm["base_mean"]) ** 2) / (2 * var) - torch.log(tm["base_scale"]) - math.log(math.sqrt(2 * math.pi))
        return lp.sum(1) + ldj
    return (-0.5 * LOG_2PI - 0.5 * z * z).sum(1) + ldj


def torch_density_step(tm, x):
    """One hot-path step as the reference executes it: C component passes, the 2-term logsumexp recursion
    (density_experiment.py:612-622), softmax weights with clamp / renormalise (:627-641).  Returns (G_ll, w)."""
    import torch
    C = tm["C"]
    rho = tm["rho"]
    G = None
    for c in range(C):
        lq = torch_component_logq(tm, x, c)
        if c == 0:
            G = lq
        else:
            r = rho[c] / rho[: c + 1].sum()
            G = torch.logsumexp(torch.stack([torch.log(1 - r) + G, torch.log(r) + lq], 1), 1)
    u = -G
    e = torch.exp(u - u.max())
    w = e / e.sum()
    if w.max() > 0.1:
        w = torch.clamp(w, 0.01, 0.1)
    if w.sum() != 1.0:
        w = w / w.sum()
    return G, w
