"""HBM roofline of the stand-alone mixture / boosting-weight pass (BASELINE config 2: POWER-shaped, N = 2 M rows, C = 4):
mixture_lse_kernel on a materialised [B, C] log q matrix, then softmax_stats / weight_apply / weight_renorm.
Algorithmic bytes: mixture 4 (C + 1) B; stats 4 B; apply 8 B; renorm 8 B."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, gbnf_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * 1024 * 1024
cfg = bench.CONFIGS["cfg2_power"]
C = cfg["C"]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = gbnf_b200.BoostedFlow(bench.make_args(cfg, dev), gemm_mode="fp32").to(dev)
model.eval()
model.pack_all()
# several independent [B, C] buffers so that successive iterations do not hit L2 (126 MB): 8 x 32 MB
NB = max(2, min(8, (1 << 30) // (B * C * 4)))   # enough buffers to exceed the 126 MB L2
bufs = [torch.randn((B, C), device=dev) * 3 - 8 for _ in range(NB)]
peak = bench.peaks()["hbm"]

def timed(fn, n=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

Gs = [model.mixture_from_logq(b, C) for b in bufs]
t_mix = timed(lambda i: model.mixture_from_logq(bufs[i % NB], C))
t_w = timed(lambda i: model.boosting_weights(Gs[i % NB]))
bytes_mix = 4 * (C + 1) * B
bytes_w = (4 + 8 + 8) * B
out = {"rows": B, "C": C,
       "mixture_lse": {"ms": t_mix, "GBps": bytes_mix / t_mix / 1e6, "frac_of_hbm_peak": bytes_mix / t_mix / 1e6 / peak},
       "boost_weights(3 kernels)": {"ms": t_w, "GBps": bytes_w / t_w / 1e6, "frac_of_hbm_peak": bytes_w / t_w / 1e6 / peak},
       "hbm_peak_GBps": peak,
       "note": "G_ll / w buffers of 8 MB stay in L2 between the three weight kernels, so their HBM traffic is below the algorithmic bytes"}
print(json.dumps(out))
