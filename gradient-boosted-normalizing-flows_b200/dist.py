"""Multi-GPU sharding of the density path: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

The reference is single-process (no distributed code at all), so this layer is new design:

* batch-parallel  -- rows are independent through log q_c and the mixture, so each rank evaluates a contiguous row
  shard with replicated parameters and NO data-path collective.  Only the batch-coupled boosting weights need
  cross-rank scalars: all-reduce(max), all-reduce(sum exp) and all-reduce(sum w) of ONE number each, between the
  three stages of the weight kernels (gbnf_weight_stats / _apply / _renorm).
* component-parallel -- rank g owns a contiguous block of components, every rank sees all rows and produces
  logq[B, C/G]; one all-gather of that block, then the local mixture kernel.

The per-rank compute goes through an `ops` object so that the collective algebra can be exercised on CPU with the
gloo backend (tests/test_dist_gloo.py substitutes the oracle there); the product default is KernelOps = the C ABI.

Data path of the product (KernelOps after `init_peer_exchange`): NO NCCL call per step.  The library owns the exchange
(include/gbnf.h, gbnf_comm_*): every rank maps every other rank's exchange block (cudaIpc, peer memory over NVLink); the
weight kernels publish / consume the two batch reductions inside the kernels themselves, and the coupling kernel's epilogue
stores log q straight into every rank's gather buffer.  torch.distributed only carries the one-time handle exchange and the
barriers around set-up / tear-down.  Without `init_peer_exchange` the same functions fall back to torch.distributed
collectives between the staged kernels (three scalar all-reduces / one all-gather) -- the algebra the gloo tests cover.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_rows(n_rows, world, rank):
    """Contiguous near-equal row shard [start, stop) of rank `rank`."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_components(n_comp, world, rank):
    """Contiguous block of component ids owned by `rank` (balanced when n_comp % world == 0)."""
    start, stop = shard_rows(n_comp, world, rank)
    return list(range(start, stop))


def init_peer_exchange(model, max_rows, group=None):
    """One-time set-up of the library-owned exchange: allocate this rank's block (+ gather buffers for `max_rows` rows),
    all-gather the 64-byte cudaIpc handles, map the peers (gbnf_comm_init), barrier."""
    world, rank = _world(group)
    lib = _lib.load()
    dev = model.rho.device
    h = model.handle(dev)
    mine = (C.c_ubyte * 64)()
    _lib.check(lib.gbnf_comm_local_handle(h, int(max_rows), mine))
    t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(allh, t, group=group)
    else:
        allh.copy_(t)
    arr = (C.c_ubyte * (64 * world))(*allh.cpu().tolist())
    _lib.check(lib.gbnf_comm_init(h, rank, world, arr))
    if world > 1:
        dist.barrier(group=group)
    return True


def shutdown_peer_exchange(model, group=None):
    """Barrier (peers must have stopped writing into this rank's block), then unmap and free."""
    world, _ = _world(group)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier(group=group)
    if model._handle is not None:
        _lib.check(_lib.load().gbnf_comm_destroy(model._handle))


class KernelOps:
    """Per-rank compute through libgbnf_b200.so (device tensors in, device tensors out).  peer=True: the library's own
    peer-memory exchange is initialised (init_peer_exchange) and the *_dist entry points are used."""

    def __init__(self, model, peer=False):
        self.model = model
        self.lib = _lib.load()
        self.peer = peer

    def boost_weights_dist(self, G_ll, mode, batch_size):
        w = torch.empty_like(G_ll)
        lo = 0.01 if mode == "density" else 0.1 / float(batch_size)
        _lib.check(self.lib.gbnf_boost_weights_dist(self._h(G_ll), self._p(G_ll), G_ll.shape[0], lo, 0.1, _lib.WEIGHTS[mode],
                                                    self._p(w), None, self._s(G_ll)))
        return w

    def mixture_component_parallel(self, x, n_comp, skip_c=-1):
        world, rank = _world(None)
        per = n_comp // max(world, 1)
        for c in range(rank * per, (rank + 1) * per):     # only the components this rank evaluates (the staleness check of all
            self.model.pack_component(c)                  # C components cost 2 ms of host time per step at cfg4)
        G = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        rho = self.model.rho.detach().to(x.device, torch.float32).contiguous()
        _lib.check(self.lib.gbnf_mixture_component_parallel(self._h(x), self._p(x), x.shape[0], n_comp, self._p(rho), skip_c,
                                                            _lib.MIX_SIMPLEX, self._p(G), self._s(x)))
        return G

    def _h(self, t):
        return self.model.handle(t.device)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    @staticmethod
    def _s(t):
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def weight_stats(self, G_ll):
        ms = torch.empty(2, device=G_ll.device, dtype=torch.float32)
        _lib.check(self.lib.gbnf_weight_stats(self._h(G_ll), self._p(G_ll), G_ll.shape[0], self._p(ms), self._s(G_ll)))
        return ms

    def weight_apply(self, G_ll, ms, lo, hi, mode):
        w = torch.empty_like(G_ll)
        wsum = torch.zeros(1, device=G_ll.device, dtype=torch.float64)
        _lib.check(self.lib.gbnf_weight_apply(self._h(G_ll), self._p(G_ll), G_ll.shape[0], self._p(ms), lo, hi,
                                              _lib.WEIGHTS[mode], self._p(w), self._p(wsum), self._s(G_ll)))
        return w, wsum

    def weight_renorm(self, w, wsum, mode):
        _lib.check(self.lib.gbnf_weight_renorm(self._h(w), self._p(w), w.shape[0], self._p(wsum), _lib.WEIGHTS[mode],
                                               self._s(w)))
        return w

    def component_logq(self, x, c0, c1):
        return self.model.component_log_density(x, c0, c1)

    def mixture(self, logq, n_comp, skip_c=-1):
        return self.model.mixture_from_logq(logq, n_comp, skip_c)

    def resample(self, w, u):
        return self.model.resample(w, u)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def merge_softmax_stats(ms, group=None):
    """Global (max, sum exp) from per-shard pairs: M = max_g M_g, S = sum_g S_g * exp(M_g - M).  Two scalar
    all-reduces.  A shard with no rows contributes (-inf, 0)."""
    world, _ = _world(group)
    if world == 1:
        return ms
    m_loc, s_loc = ms[0:1].clone(), ms[1:2].clone()
    m_glob = m_loc.clone()
    dist.all_reduce(m_glob, op=dist.ReduceOp.MAX, group=group)
    scaled = torch.where(torch.isfinite(m_loc), s_loc * torch.exp(m_loc - m_glob), torch.zeros_like(s_loc))
    dist.all_reduce(scaled, op=dist.ReduceOp.SUM, group=group)
    return torch.cat([m_glob, scaled])


def boosting_weights_batch_parallel(ops, G_ll_local, mode="density", batch_size=None, group=None):
    """Weights of this rank's row shard under the GLOBAL batch softmax (density_experiment.py:627-641 semantics on the
    union of all shards).  `batch_size` (toy clamp floor 0.1 / batch_size) is the global batch size."""
    n_local = G_ll_local.shape[0]
    if getattr(ops, "peer", False):      # library-owned exchange through peer memory: no collective call on the data path
        return ops.boost_weights_dist(G_ll_local, mode, batch_size)
    if n_local > 0:
        ms = ops.weight_stats(G_ll_local)
    else:
        ms = torch.tensor([float("-inf"), 0.0], device=G_ll_local.device, dtype=torch.float32)
    ms = merge_softmax_stats(ms, group)
    lo = 0.01 if mode == "density" else 0.1 / float(batch_size)
    if n_local > 0:
        w, wsum = ops.weight_apply(G_ll_local, ms, lo, 0.1, mode)
    else:
        w, wsum = G_ll_local.clone(), torch.zeros(1, device=G_ll_local.device, dtype=torch.float64)
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(wsum, op=dist.ReduceOp.SUM, group=group)
    if n_local > 0:
        w = ops.weight_renorm(w, wsum, mode)
    return w


def mixture_component_parallel(ops, x, n_comp, skip_c=-1, group=None):
    """G_ll[B] with the components split across ranks: local logq block -> all-gather -> local logsumexp.
    Every rank holds all rows of x; n_comp must be divisible by the world size for the single all-gather."""
    world, rank = _world(group)
    if world == 1:
        return ops.mixture(ops.component_logq(x, 0, n_comp), n_comp, skip_c)
    if getattr(ops, "peer", False):      # coupling epilogue stores log q into every rank's gather buffer (NVLink peer memory)
        return ops.mixture_component_parallel(x, n_comp, skip_c)
    if n_comp % world != 0:
        raise ValueError("component-parallel evaluation needs n_comp % world_size == 0")
    per = n_comp // world
    local = ops.component_logq(x, rank * per, (rank + 1) * per).contiguous()       # [B, per]
    gathered = torch.empty((world * local.shape[0], per), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(gathered, local, group=group)                       # rank-major [world * B, per]
    logq = gathered.view(world, local.shape[0], per).permute(1, 0, 2).reshape(local.shape[0], n_comp).contiguous()
    return ops.mixture(logq, n_comp, skip_c)


def resample_batch_parallel(ops, w_local, x_local, u_global, group=None):
    """Global inverse-CDF resampling when rows are sharded: weights and rows are all-gathered (batch-sized, small),
    every rank resolves ITS slice of the shared uniforms against the global CDF and returns its resampled rows.
    All shards must have equal length (pad upstream)."""
    world, rank = _world(group)
    if world == 1:
        idx = ops.resample(w_local, u_global)
        return idx, x_local[idx]
    w_all = torch.empty(world * w_local.shape[0], device=w_local.device, dtype=w_local.dtype)
    dist.all_gather_into_tensor(w_all, w_local.contiguous(), group=group)
    x_all = torch.empty((world * x_local.shape[0], x_local.shape[1]), device=x_local.device, dtype=x_local.dtype)
    dist.all_gather_into_tensor(x_all, x_local.contiguous(), group=group)
    lo, hi = shard_rows(u_global.shape[0], world, rank)
    idx = ops.resample(w_all, u_global[lo:hi].contiguous())
    return idx, x_all[idx]
