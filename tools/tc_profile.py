"""Cycle breakdown of the tensor-core coupling kernel (CTA 0) on the benchmark configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, gbnf_b200

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_miniboone"
mode = sys.argv[2] if len(sys.argv) > 2 else "f16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
cfg = bench.CONFIGS[cfg_name]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
margs = bench.make_args(cfg, dev)
margs.coupling_network = os.environ.get('GBNF_PROF_ACT', margs.coupling_network)   # relu: MUFU-free epilogue (isolates MMA + weight stream)
model = gbnf_b200.BoostedFlow(margs, gemm_mode=mode).to(dev)
x = torch.randn((B, cfg["D"]), device=dev)
model.train()
with torch.no_grad():
    for c in range(cfg["C"]):
        model(x=x[:4096], components=c)
model.eval()
for p in model.parameters():
    p.requires_grad_(False)
model.pack_all()
for _ in range(300):      # SM clocks ramp up over the first few hundred ms: cycle counters of a cold launch are not representative
    G = model.mixture_log_density(x, cfg["C"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    G = model.mixture_log_density(x, cfg["C"])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
p = model.profile()
tiles0 = (B + 127) // 128
tiles_cta0 = (tiles0 + 147) // 148 if tiles0 > 148 else 1
steps = tiles_cta0 * cfg["C"] * cfg["K"]
print(f"{cfg_name} {mode} B={B}: {ms:.3f} ms, {B / ms * 1e-3:.2f} M samples/s, flops {bench.flops_per_sample(cfg) * B / ms / 1e9:.1f} TF")
print(f"CTA0: tile-steps {steps}; per tile-step cycles:")
print(f"  MMA warp  : total {p[0] / steps:9.0f}  wait a_ready {p[1] / steps:9.0f}  wait weights {p[2] / steps:9.0f}  other {(p[0] - p[1] - p[2]) / steps:9.0f}  layers {p[3]}")
print(f"  MMA warp+ : issue blocks {p[4] / steps:9.0f}  wait sr {p[5] / steps:9.0f}")
print(f"  epilogue  : total {p[8] / steps:9.0f}  wait acc {p[9] / steps:9.0f}  hidden {p[10] / steps:9.0f}  last {p[11] / steps:9.0f}  prologue {p[12] / steps:9.0f}")
print(f"  epilogue+ : wait L1 {p[13] / steps:7.0f}  wait L2 {p[14] / steps:7.0f}  wait L3 {p[15] / steps:7.0f}  L1 epi {p[19] / steps:7.0f}  L2 epi {p[20] / steps:7.0f}  x-load {p[21] / steps:7.0f}  gather {p[22] / steps:7.0f}  pipelined={model.info()["pipelined"]}")
print(f"  producer  : total {p[16] / steps:9.0f}  wait empty {p[17] / steps:9.0f}  stages {p[18]}")
