"""cfg5 (h = 1024, CTA-pair kernel): launch geometry and timing at a few batch sizes (diagnostics)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, gbnf_b200
cfg = bench.CONFIGS["cfg5_bsds300"]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = gbnf_b200.BoostedFlow(bench.make_args(cfg, dev), gemm_mode=sys.argv[1] if len(sys.argv) > 1 else "f16fast").to(dev)
x = torch.randn((1 << 17, cfg["D"]), device=dev)
model.train()
with torch.no_grad():
    for c in range(cfg["C"]):
        model(x=x[:4096], components=c)
model.eval()
for p in model.parameters():
    p.requires_grad_(False)
model.pack_all()
for B in (128 * 74, 128 * 74 * 2, 65536, 1 << 17):
    xb = x[:B].contiguous()
    for _ in range(2):
        model.mixture_log_density(xb, cfg["C"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        model.mixture_log_density(xb, cfg["C"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    inf = model.info()
    tf = bench.flops_per_sample(cfg) * B / (ms * 1e-3) / 1e12
    print(f"B={B}: {ms:.3f} ms  {B / ms * 1e3 / 1e6:.3f} M samples/s  {tf:.0f} TF  grid {inf['grid']} smem {inf['smem_bytes']}", flush=True)
if os.environ.get("GBNF_PROF"):
    pr = model.profile()
    names = {0: "mma total", 1: "mma wait ring", 2: "mma wait a0r/a1r", 3: "mma wait slotB", 4: "mma wait piece", 5: "passes", 8: "epi total", 9: "epi wait l1f",
             10: "epi wait l2own", 11: "epi wait xfull", 12: "epi wait l2ship", 13: "epi wait xfree", 14: "epi wait l3f", 15: "epi wait x3full",
             16: "prod total", 17: "prod wait empty", 19: "epi gather", 20: "epi l1 work", 21: "epi own work", 22: "epi ship work", 23: "epi transform+bar"}
    n = max(pr[5], 1)
    for i, nm in names.items():
        print(f"  {nm:22s} {pr[i]:>12d}  per pass {pr[i] / n:10.0f}")
