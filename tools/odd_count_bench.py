"""Fixed mixture with an ODD number of components (what the training loop evaluates for component index 7, 5, 3): two-chain kernel on
the even part + single-chain kernel on the last component vs the single-chain kernel on all of them (GBNF_TC4=0)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, gbnf_b200

cfg = bench.CONFIGS["cfg3_miniboone"]
dev = torch.device("cuda", 0)
out = {}
for tag, env in (("two_chain_plus_last", None), ("single_chain", "0")):
    if env is None:
        os.environ.pop("GBNF_TC4", None)
    else:
        os.environ["GBNF_TC4"] = env
    torch.manual_seed(1)
    model = gbnf_b200.BoostedFlow(bench.make_args(cfg, dev), gemm_mode="f16fast").to(dev)
    x = torch.randn((65536, cfg["D"]), device=dev)
    model.train()
    with torch.no_grad():
        for c in range(cfg["C"]):
            model(x=x[:4096], components=c)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    model.pack_all()
    for n in (7, 5, 3):
        for _ in range(20):
            model.mixture_log_density(x, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            model.mixture_log_density(x, n)
        e1.record(); torch.cuda.synchronize()
        out[f"{tag}.n{n}.ms"] = e0.elapsed_time(e1) / 100
    model.release()
print(json.dumps(out))
