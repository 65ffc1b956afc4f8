"""Multi-GPU plumbing: one process per GPU, particles / chains sharded by GLOBAL index, one all-reduce of the
(weighted) count tensor.  torch.distributed is used for exactly that collective (NCCL over NVLink on the GPU
box, gloo in the CPU tests) and nothing else.

Sharding contract: rank g of G handles indices [g*N//G, (g+1)*N//G).  Because the Philox
counter is the global particle / chain index, the all-reduced tensor equals the single-GPU tensor up to
fp64 addition order (exactly, for the integer Gibbs counts).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Tuple

import numpy as np

from . import _lib
from ._lib import check, ptr
from .device_layout import device_net


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first index, count) of rank's contiguous shard; shards tile [0, n_total) exactly."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi - lo


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def allreduce_sum_(t):
    """In-place sum over ranks of a count tensor (torch tensor on this rank's device, or CPU for gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allgather_rows_(local, counts):
    """Concatenate per-rank row blocks (rank g holds counts[g] rows of equal width) on every rank.  Rows are padded to
    the largest block so that one all_gather suffices; returns the [sum(counts), width] tensor."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return local
    most = max(counts)
    pad = local.new_zeros((most,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


class ShardedLikelihoodWeighting:
    """LW inference whose particles are split over the ranks of the default process group.

    step() enqueues this rank's shard (bn_lw_infer_dev) on torch's current stream into a resident fp64
    tensor [ncell + 2] = (weighted counts | sum w | sum w^2) and all-reduces it; normalisation happens after
    the reduction, exactly where the reference normalises (src/Inference/likelihood.jl:56)."""

    def __init__(self, bn, query, evidence, n_total: int, seed: int = 0, device: int | None = None):
        import torch
        self.torch = torch
        rank, world, local = env_rank_world()
        self.rank, self.world = rank, world
        self.device = local if device is None else device
        self.net = device_net(bn, self.device)
        lay = self.net.layout
        self.ev = lay.evidence_vector(evidence)
        self.q = lay.query_indices(query)
        self.shape = [int(lay.card[i]) for i in self.q]
        self.ncell = int(np.prod(self.shape)) if self.shape else 1
        self.n_total = int(n_total)
        self.first, self.count = shard_range(self.n_total, rank, world)
        self.seed = int(seed)
        self.out = torch.zeros(self.ncell + 2, dtype=torch.float64, device=f"cuda:{self.device}")

    def launch(self, first=None, count=None):
        """Enqueue the local shard only (no collective); result in self.out."""
        torch = self.torch
        _lib.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        qp = self.q if self.q.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_lw_infer_dev(_lib.h64(self.net.handle), ptr(self.ev, _lib.i8p), ptr(qp, _lib.i32p),
                                         C.c_int32(self.q.size),
                                         C.c_uint64(self.first if first is None else first),
                                         C.c_uint64(self.count if count is None else count),
                                         C.c_uint64(self.seed), C.c_void_p(self.out.data_ptr())))
        return self.out

    def step(self):
        self.launch()
        return allreduce_sum_(self.out)

    def posterior(self) -> np.ndarray:
        c = self.out[:self.ncell].cpu().numpy()
        with np.errstate(invalid="ignore", divide="ignore"):
            return (c / c.sum()).reshape(self.shape, order="F")


class ShardedGibbs:
    """Gibbs inference (Nodewise / Full) whose CHAINS are split over the ranks of the default process group
    (BASELINE config C4: 65 536 chains over 8 GPUs).  Rank g runs chains [g*C//G, (g+1)*C//G) of one global chain
    index space (the Philox counter is the global chain index), then the integer count tensors are all-reduced:
    the result equals the single-GPU counts exactly."""

    def __init__(self, bn, query, evidence, n_chains: int, nsamples: int, burn_in: int, thin: int, full: bool = False,
                 seed: int = 0, device: int | None = None):
        import torch
        self.torch = torch
        rank, world, local = env_rank_world()
        self.rank, self.world = rank, world
        self.device = local if device is None else device
        self.net = device_net(bn, self.device)
        lay = self.net.layout
        self.ev = lay.evidence_vector(evidence)
        self.q = lay.query_indices(query)
        self.shape = [int(lay.card[i]) for i in self.q]
        self.ncell = int(np.prod(self.shape)) if self.shape else 1
        self.first, self.count = shard_range(int(n_chains), rank, world)
        self.nsamples, self.burn_in, self.thin, self.full, self.seed = int(nsamples), int(burn_in), int(thin), bool(full), int(seed)
        self.out = torch.zeros(self.ncell, dtype=torch.float64, device=f"cuda:{self.device}")

    def launch(self):
        """Enqueue the local shard only (no collective); counts in self.out."""
        _lib.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)
        qp = self.q if self.q.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_gibbs_infer_dev(_lib.h64(self.net.handle), ptr(self.ev, _lib.i8p), ptr(qp, _lib.i32p),
                                            C.c_int32(self.q.size), C.c_uint64(self.first), C.c_uint64(self.count),
                                            C.c_int64(self.nsamples), C.c_int64(self.burn_in), C.c_int64(self.thin),
                                            C.c_int32(1 if self.full else 0), C.c_uint64(self.seed),
                                            C.c_void_p(self.out.data_ptr())))
        return self.out

    def step(self):
        self.launch()
        return allreduce_sum_(self.out)

    def posterior(self) -> np.ndarray:
        c = self.out.cpu().numpy()
        with np.errstate(invalid="ignore", divide="ignore"):
            return (c / c.sum()).reshape(self.shape, order="F")


class ShardedTableStatistics:
    """The learn-from-samples loop with the table's ROWS split over the ranks of the default process group:
    rank g draws rows [g*R//G, (g+1)*R//G) of one global row index space (the Philox counter is the global row index,
    so the union of the shards IS the single-GPU table), counts its shard where it lies (bn_statistics_dev) and the
    int64 count tensors are all-reduced -- integers, so every rank holds exactly the single-GPU counts.  The data
    log-likelihood is then the counts dotted with the log CPTs (bn_logpdf_counts_dev): the same bits on every rank and
    for every number of ranks.  No data-path collective besides that one all-reduce of cpt_len int64."""

    def __init__(self, bn, n_rows_total: int, seed: int = 0, device: int | None = None):
        import torch
        self.torch = torch
        rank, world, local = env_rank_world()
        self.rank, self.world = rank, world
        self.device = local if device is None else device
        self.net = device_net(bn, self.device)
        lay = self.net.layout
        self.n_rows_total = int(n_rows_total)
        self.first, self.count = shard_range(self.n_rows_total, rank, world)
        self.seed = int(seed)
        dev = f"cuda:{self.device}"
        self.table = torch.empty((lay.n, max(self.count, 1)), dtype=torch.uint8, device=dev)  # pitch = row count
        self.counts = torch.zeros(int(lay.cpt_off[-1]), dtype=torch.int64, device=dev)
        self.logl = torch.zeros(1, dtype=torch.float64, device=dev)

    def sample(self):
        """rand(bn, n): this rank's rows, left in HBM (node-major uint8, one column of `count` bytes per node)."""
        _lib.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)
        if self.count:
            check(_lib.lib().bn_sample_dev(_lib.h64(self.net.handle), None, C.c_int32(0), C.c_int32(1), C.c_uint64(self.first),
                                           C.c_uint64(self.count), C.c_uint64(self.seed), C.c_void_p(self.table.data_ptr()),
                                           None, None))
        return self.table

    def launch(self):
        """Count the local shard only (no collective); int64 counts in self.counts."""
        _lib.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)
        check(_lib.lib().bn_statistics_dev(_lib.h64(self.net.handle), C.c_void_p(self.table.data_ptr()), C.c_uint64(self.count),
                                           C.c_uint64(max(self.count, 1)), C.c_void_p(self.counts.data_ptr())))
        return self.counts

    def step(self):
        """sample -> count -> all-reduce: the whole table's sufficient statistics on every rank."""
        self.sample()
        self.launch()
        return allreduce_sum_(self.counts)

    def logpdf(self) -> float:
        """logpdf(bn, table) of the WHOLE table from the all-reduced counts (call after step())."""
        _lib.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)
        check(_lib.lib().bn_logpdf_counts_dev(_lib.h64(self.net.handle), C.c_void_p(self.counts.data_ptr()),
                                              C.c_void_p(self.logl.data_ptr())))
        return float(self.logl.cpu()[0])


class ShardedLoopy:
    """LoopyBelief over many (query, evidence) pairs, the PAIRS split over the ranks of the default process group:
    rank g answers pairs [g*B//G, (g+1)*B//G) with one bn_loopy_infer_dev launch; pairs are independent, so there is no
    data-path collective (weak scaling) -- gather() all-gathers the beliefs only if every rank wants all of them."""

    def __init__(self, bn, queries, evidences, nsamples: int = 500, tol: float = 1e-8, iters_for_convergence: int = 6,
                 device: int | None = None):
        import torch
        self.torch = torch
        rank, world, local = env_rank_world()
        self.rank, self.world = rank, world
        self.device = local if device is None else device
        self.net = device_net(bn, self.device)
        lay = self.net.layout
        self.n_total = len(queries)
        self.first, self.count = shard_range(self.n_total, rank, world)
        self.counts = [shard_range(self.n_total, g, world)[1] for g in range(world)]
        mine = range(self.first, self.first + self.count)
        ev = np.stack([lay.evidence_vector(evidences[b]) for b in mine]) if self.count else np.zeros((0, lay.n), np.int8)
        q = np.array([lay.index[queries[b]] for b in mine], np.int32)
        dev = f"cuda:{self.device}"
        self.stride = int(lay.card.max())
        self.ev_d = torch.from_numpy(ev).to(dev)
        self.q_d = torch.from_numpy(q).to(dev)
        self.out = torch.zeros((self.count, self.stride), dtype=torch.float64, device=dev)
        self.iters = torch.zeros(self.count, dtype=torch.int32, device=dev)
        self.nsamples, self.tol, self.K = int(nsamples), float(tol), int(iters_for_convergence)

    def launch(self):
        """Enqueue the local shard on torch's current stream; beliefs in self.out [count, max_card]."""
        _lib.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)
        check(_lib.lib().bn_loopy_infer_dev(_lib.h64(self.net.handle), C.c_void_p(self.ev_d.data_ptr()),
                                            C.c_void_p(self.q_d.data_ptr()), C.c_uint64(self.count), C.c_int32(self.nsamples),
                                            C.c_double(self.tol), C.c_int32(self.K), C.c_int32(self.stride),
                                            C.c_void_p(self.out.data_ptr()), C.c_void_p(self.iters.data_ptr())))
        return self.out

    def gather(self):
        """All pairs' beliefs on every rank, in pair order."""
        return allgather_rows_(self.out, self.counts)
