"""Timeline (clock64 stamps, CTA 0, one coupling pass) of the pipelined tensor-core kernel.  Needs GBNF_PROF=1."""
import os, sys
os.environ.setdefault("GBNF_PROF", "2")   # 2: event trace only (near-production timing), 1: plus cycle counters
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, gbnf_b200

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_miniboone"
mode = sys.argv[2] if len(sys.argv) > 2 else "f16fast"
cfg = bench.CONFIGS[cfg_name]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
margs = bench.make_args(cfg, dev)
margs.coupling_network = os.environ.get('GBNF_PROF_ACT', margs.coupling_network)
model = gbnf_b200.BoostedFlow(margs, gemm_mode=mode).to(dev)
x = torch.randn((65536, cfg["D"]), device=dev)
model.train()
with torch.no_grad():
    for c in range(cfg["C"]):
        model(x=x[:4096], components=c)
model.eval()
for p in model.parameters():
    p.requires_grad_(False)
model.pack_all()
for _ in range(30):   # let the clocks ramp up; the trace of the last launch is read
    model.mixture_log_density(x, cfg["C"])
torch.cuda.synchronize()
t = model.trace()
NQ, NJ = cfg["h"] // 128, cfg["h"] // 64
names = {0: "MMA  a0r seen"}
for q in range(NQ):
    names[1 + q] = f"MMA  L1({q}) issued"; names[5 + q] = f"MMA  a1r[{q}] seen"
for j in range(NJ):
    names[10 + j] = f"MMA  L2 chunk {j} issued+committed"; names[20 + j] = f"MMA  L3({j}) issued"
for g in (0, 1):
    b = 40 + 40 * g
    names[b] = f"EPI{g} a0r arrived (gather done)"
    for q in range(NQ):
        names[b + 1 + q] = f"EPI{g} l1f[{q}] seen"; names[b + 5 + q] = f"EPI{g} a1r[{q}] arrived (L1 chunk packed)"
    for j in range(NJ):
        names[b + 10 + j] = f"EPI{g} l2f chunk {j} seen"; names[b + 20 + j] = f"EPI{g} sr chunk {j} arrived"
    names[b + 30] = f"EPI{g} l3f seen"; names[b + 31] = f"EPI{g} last-layer epilogue done"
    names[b + 32] = f"EPI{g} pass top"; names[b + 33] = f"EPI{g} gather loop done"
    names[b + 34] = f"EPI{g} end-of-pass barrier passed"
names[201] = "EPI0 NEXT pass top"; names[200] = "EPI0 NEXT pass a0r arrived"
for i, n in enumerate(["L1 chunk 1: tcgen05.ld done", "L1 chunk 1: bias+act+pack done", "L1 chunk 1: tcgen05.st done"]):
    names[120 + i] = "EPI0-w0 " + n
for i, n in enumerate(["L2(4,h0) op start", "L2(4,h0) stage acquired", "L2(4,h0) 16 MMA issued", "L2(4,h1) op start", "L2(4,h1) stage acquired", "L2(4,h1) 16 MMA issued", "L3(1) op start", "L3(1) piece barrier passed"]):
    names[130 + i] = "MMA* " + n
names[138] = "MMA* layer-1 weight stage taken"
ev = sorted((v, names[i]) for i, v in enumerate(t) if v > 0 and i in names)
t0 = ev[0][0]
print(f"{cfg_name} {mode}: one coupling pass of CTA 0 (cycles relative to first event)")
if t[200] > 0 and t[40] > 0:
    print(f"PASS LENGTH (a0r arrival of this pass -> a0r arrival of the next): {t[200] - t[40]} cycles")
for v, n in ev:
    print(f"{v - t0:8d}  {n}")
