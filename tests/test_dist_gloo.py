"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in gbnf_b200.dist.  The per-rank compute is supplied by
the CPU oracle here (an `ops` object with the same methods as dist.KernelOps); what is under test is the sharding and
the collective algebra: merged softmax statistics, global clamp decision, global renormalisation, component gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleOps:
    def __init__(self, md):
        from oracle import gbnf_oracle as orc
        self.orc, self.md = orc, md

    def weight_stats(self, G):
        u = -G.numpy()
        m = u.max()
        return torch.tensor([m, np.exp(u - m).sum()], dtype=torch.float32)

    def weight_apply(self, G, ms, lo, hi, mode):
        w = np.exp(-G.numpy() - ms[0].item()) / np.float32(ms[1].item())
        if np.float32(1.0) / np.float32(ms[1].item()) > hi:
            w = np.clip(w, np.float32(lo), np.float32(hi))
        w = w.astype(np.float32)
        return torch.from_numpy(w), torch.tensor([w.astype(np.float64).sum()], dtype=torch.float64)

    def weight_renorm(self, w, wsum, mode):
        s = np.float32(wsum.item())
        return w if (mode == "density" and s == 1.0) else w / s

    def component_logq(self, x, c0, c1):
        return torch.from_numpy(np.stack([self.orc.component_logq(self.md, x.numpy(), c) for c in range(c0, c1)], 1))

    def mixture(self, logq, n, skip_c=-1):
        return torch.from_numpy(self.orc.mixture_flat(logq.numpy(), self.md["rho"], n, skip_c))

    def resample(self, w, u):
        return torch.from_numpy(self.orc.resample_indices(w.numpy(), u.numpy()))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gbnf_b200 import dist as gd
        from oracle import gbnf_oracle as orc
        md = orc.make_synthetic_model("glow", 6, 4, 2, 16, seed=5, init_rows=128)
        ops = OracleOps(md)
        rng = np.random.default_rng(11)
        x = rng.standard_normal((301, 6)).astype(np.float32)          # odd size: ragged shards
        res = {}
        for scale, tag in ((1.0, "clamped"), (0.01, "flat")):
            G = orc.mixture_flat(orc.all_component_logq(md, x), md["rho"], 4) * np.float32(scale)
            lo, hi = gd.shard_rows(len(G), world, rank)
            w = gd.boosting_weights_batch_parallel(ops, torch.from_numpy(G[lo:hi].copy()), "density")
            res["w_" + tag] = (lo, hi, w.numpy(), orc.boost_weights(G, "density")[lo:hi])
            wt = gd.boosting_weights_batch_parallel(ops, torch.from_numpy(G[lo:hi].copy()), "toy", batch_size=len(G))
            res["wt_" + tag] = (lo, hi, wt.numpy(), orc.boost_weights(G, "toy", len(G))[lo:hi])
        Gcp = gd.mixture_component_parallel(ops, torch.from_numpy(x), 4)
        res["cp"] = (Gcp.numpy(), orc.mixture_flat(orc.all_component_logq(md, x), md["rho"], 4))
        # resampling with sharded rows (equal shards): 300 rows
        xs = x[:300]; lo, hi = gd.shard_rows(300, world, rank)
        wfull = orc.boost_weights(orc.mixture_flat(orc.all_component_logq(md, xs), md["rho"], 4), "density")
        u = np.random.default_rng(3).random(300)
        idx, xr = gd.resample_batch_parallel(ops, torch.from_numpy(wfull[lo:hi].copy()), torch.from_numpy(xs[lo:hi].copy()),
                                             torch.from_numpy(u))
        full_idx = orc.resample_indices(wfull, u)
        res["rs"] = (idx.numpy(), full_idx[lo:hi], xr.numpy(), xs[full_idx[lo:hi]])
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    covered = []
    for r in range(world):
        res = results[r]
        for k in ("w_clamped", "w_flat", "wt_clamped", "wt_flat"):
            lo, hi, got, want = res[k]
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-10, err_msg=k)
        covered.append((res["w_flat"][0], res["w_flat"][1]))
        np.testing.assert_allclose(res["cp"][0], res["cp"][1], rtol=1e-6, atol=1e-6)
        assert np.array_equal(res["rs"][0], res["rs"][1])
        np.testing.assert_array_equal(res["rs"][2], res["rs"][3])
    assert covered == [(0, 151), (151, 301)]


def test_shard_helpers():
    from gbnf_b200 import dist as gd
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [gd.shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert gd.shard_components(16, 8, 3) == [6, 7]
