"""Error distribution of the f16 tensor-core modes against the fp64 oracle at the bench batch (diagnostics)."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_model
from oracle import gbnf_oracle as orc
CFG = {"cfg3": dict(kind="glow", D=43, C=8, K=5, h=512), "cfg4": dict(kind="glow", D=21, C=16, K=10, h=512)}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for name, kw in CFG.items():
    kw = dict(kw)
    md = orc.make_synthetic_model(kw.pop("kind"), kw.pop("D"), kw.pop("C"), kw.pop("K"), kw.pop("h"), seed=1)
    x = np.random.default_rng(1234).standard_normal((B, md["D"])).astype(np.float32)
    m64 = orc.cast_model(md, np.float64)
    ref = np.concatenate([orc.all_component_logq(m64, x[s:s + 8192].astype(np.float64)) for s in range(0, B, 8192)], 0)
    for mode in ("fp32", "f16", "f16fast"):
        model = build_model(md, "cuda", gemm_mode=mode)
        Bm = B if mode != "fp32" else min(B, 4096)
        lq = model.component_log_density(torch.from_numpy(x[:Bm]).cuda()).cpu().numpy().astype(np.float64)
        r = ref[:Bm]
        rel = np.abs(lq - r) / np.abs(r)
        ab = np.abs(lq - r)
        print(f"{name} {mode} B={Bm}: |logq| min {np.abs(r).min():.2f} med {np.median(np.abs(r)):.1f}; rel max {rel.max():.2e} p99.9 {np.quantile(rel, 0.999):.2e} "
              f"p50 {np.median(rel):.2e}; abs max {ab.max():.2e} p99.9 {np.quantile(ab, 0.999):.2e}; n(rel>1e-4) {(rel > 1e-4).sum()} of {rel.size}; "
              f"worst at |logq| {np.abs(r).flat[rel.argmax()]:.2f}", flush=True)
        model.release()
