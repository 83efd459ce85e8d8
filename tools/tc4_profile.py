"""Per-role cycle counters of CTA 0 of the two-chain kernel (csrc/coupling_tc4.cuh) on a benchmark configuration.
usage: GBNF_PROF=1 python tools/tc4_profile.py [cfg] [mode] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GBNF_PROF", "1")
import torch
import bench, gbnf_b200

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_miniboone"
mode = sys.argv[2] if len(sys.argv) > 2 else "f16fast"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
cfg = bench.CONFIGS[cfg_name]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = gbnf_b200.BoostedFlow(bench.make_args(cfg, dev), gemm_mode=mode).to(dev)
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
inf = model.info()
n = max(1, p[6])
print(f"{cfg_name} {mode} B={B}: {ms:.3f} ms  {B / ms * 1e-3:.2f} M samples/s  two_chain={inf['two_chain']}  passes of CTA 0: {p[6]}")
print(f"per pass (cycles):")
print(f"  MMA warp : total {p[0] / n:7.0f} | wait a0r {p[1] / n:6.0f}  ring {p[2] / n:6.0f}  sr+w3 {p[3] / n:6.0f}  l3d {p[4] / n:6.0f}  a1r {p[5] / n:6.0f}")
print(f"  epilogue : total {p[8] / n:7.0f} | wait l1f {p[9] / n:6.0f}  l2f {p[10] / n:6.0f}  l3f {p[11] / n:6.0f} | finish_pass {p[12] / n:6.0f} (transform {p[19] / n:5.0f} gather {p[20] / n:5.0f})  end-of-slot barrier {p[13] / n:6.0f}  E1 {p[14] / n:6.0f}  E2 {p[15] / n:6.0f}")
print(f"  producer : total {p[16] / n:7.0f} | wait empty {p[17] / n:6.0f}  wait w3empty {p[18] / n:6.0f}")
tr = model.trace()
ng = inf["grid"]
tot = [t & 0xffffffff for t in tr[:ng]]
ring = [t >> 32 for t in tr[:ng]]
import statistics
print("per-CTA MMA-warp total cycles: min %d  median %d  max %d ; all waits: min %d median %d max %d" % (min(tot), statistics.median(tot), max(tot), min(ring), statistics.median(ring), max(ring)))
print("slowest CTAs:", sorted(range(ng), key=lambda i: -tot[i])[:12])
print("totals by CTA (k cycles):", [t // 1000 for t in tot])
print("busy (total - waits) by CTA (k cycles):", [(t - r) // 1000 for t, r in zip(tot, ring)])

names = {0: "MMA slot start (ring stage + a0r ok)", 1: "MMA L1(0) issued", 2: "MMA prev L3(4) issued", 3: "MMA a1r[0] seen", 4: "MMA a1r[1] seen",
         5: "MMA a1r[2] seen", 6: "MMA L2c0 parts 1,2 issued", 7: "MMA a1r[3] seen", 8: "MMA L2c0 part 3 issued",
         10: "MMA c1: start", 11: "MMA c1: (no piece)", 12: "MMA c1 issued", 13: "MMA c2: start", 14: "MMA L3(0) issued", 15: "MMA c2 issued",
         16: "MMA c3: start", 17: "MMA L3(1) issued", 18: "MMA c3 issued", 19: "MMA c4: start", 20: "MMA L3(2) issued", 21: "MMA c4 issued",
         22: "MMA before L3(3)", 23: "MMA L3(3) issued",
         40: "EPI t0 start", 41: "EPI E2[4](A) done", 44: "EPI E1[0] done", 45: "EPI E1[1] done", 46: "EPI E1[2] done", 47: "EPI E1[3] done",
         42: "EPI transform(A) done", 48: "EPI t1 start", 49: "EPI E2[0] done", 50: "EPI t1 end", 56: "EPI t2 start", 57: "EPI E2[1] done",
         58: "EPI finish_b(A) done", 64: "EPI t3 start", 65: "EPI E2[2] done", 66: "EPI t3 end", 72: "EPI t4 start", 73: "EPI E2[3] done", 74: "EPI t4 end"}
full = model.trace()
ev = [(full[158 + i], names[i]) for i in names if full[158 + i] > 0]   # a.prof[190 + id] = trace[158 + id]
if ev:
    t0 = min(t for t, _ in ev)
    print("event trace of slot 41 of CTA 0 (cycles from the first event):")
    for t, n in sorted(ev):
        print(f"  {t - t0:7d}  {n}")
