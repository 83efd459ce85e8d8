"""Multi-GPU parity of the sharded paths (run under torchrun, one process per GPU, NCCL):
batch-parallel (row shards, global boosting weights via three scalar all-reduces, global resampling) and
component-parallel (component blocks, one all-gather of log q) must reproduce the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import build_model, rel_err  # noqa: E402
from oracle import gbnf_oracle as orc  # noqa: E402  (test infrastructure: the checker)
from gbnf_b200 import dist as gd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C = 4 * world if world <= 4 else 16
    md = orc.make_synthetic_model("glow", 21, C, 2, 256, seed=3)
    B = 128 * 3 * world + 7 * world          # equal shards, ragged tiles
    x_np = np.random.default_rng(99).standard_normal((B, 21)).astype(np.float32)
    x = torch.from_numpy(x_np).to(dev)
    for mode, peer in (("fp32", False), ("fp32", True), ("f16fast", True), ("f16fast", False)):
        model = build_model(md, dev, gemm_mode=mode)
        ops = gd.KernelOps(model, peer=peer)
        if peer:     # the library-owned peer-memory exchange (gbnf_comm_*): no NCCL call on the data path
            gd.init_peer_exchange(model, B)
        # single-GPU result (every rank computes it: replicated parameters)
        G = model.mixture_log_density(x, C)
        w = model.boosting_weights(G)
        u = torch.rand(B, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        idx = model.resample(w, u)
        # oracle on rank 0 pins the single-GPU result itself
        if rank == 0:
            ref = orc.mixture_recursion(orc.all_component_logq(orc.cast_model(md, np.float64), x_np.astype(np.float64)),
                                        md["rho"].astype(np.float64), C)
            assert rel_err(G.cpu().numpy(), ref) < (1e-5 if mode == "fp32" else 1e-4), mode
        # ---- batch-parallel ----
        lo, hi = gd.shard_rows(B, world, rank)
        xs = x[lo:hi].contiguous()
        G_loc = model.mixture_log_density(xs, C)
        assert torch.equal(G_loc, G[lo:hi]), "rows are independent: a row shard must reproduce its rows bitwise"
        w_loc = gd.boosting_weights_batch_parallel(ops, G_loc, "density")
        w_all = torch.empty(B, device=dev)
        dist.all_gather_into_tensor(w_all, w_loc.contiguous())
        assert rel_err(w_all.cpu().numpy(), w.cpu().numpy()) < 5e-6, (mode, rel_err(w_all.cpu().numpy(), w.cpu().numpy()))
        idx_loc, xr_loc = gd.resample_batch_parallel(ops, w[lo:hi].contiguous(), xs, u)
        assert torch.equal(idx_loc, idx[lo:hi]) and torch.equal(xr_loc, x[idx[lo:hi]])
        # ---- component-parallel ----
        G_cp = gd.mixture_component_parallel(ops, x, C)
        e = rel_err(G_cp.cpu().numpy(), G.cpu().numpy())
        assert e < 2e-6, (mode, e)
        if peer:     # several epochs back to back (double-buffered slots), then a different batch size
            for it in range(5):
                w2 = gd.boosting_weights_batch_parallel(ops, G_loc, "density")
                assert torch.equal(w2, w_loc), "peer exchange must be deterministic across epochs"
                G2 = gd.mixture_component_parallel(ops, x, C)
                assert torch.equal(G2, G_cp)
            w3 = gd.boosting_weights_batch_parallel(ops, G_loc[:100].contiguous(), "toy", batch_size=100 * world)
            assert abs(float(w3.double().sum()) * 1.0 - 0.0) >= 0.0
            tot = w3.double().sum().reshape(1)
            dist.all_reduce(tot)
            assert abs(float(tot) - 1.0) < 1e-5
            gd.shutdown_peer_exchange(model)
        torch.cuda.synchronize()
        model.release()
        if rank == 0:
            print(f"[{mode}{' peer-memory exchange' if peer else ' torch.distributed collectives'}] world {world}: batch-parallel weights / resampling and component-parallel mixture match (G cp rel {e:.1e})")
    dist.barrier()
    if rank == 0:
        print("MULTI-GPU CHECK OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
