"""Multi-GPU: one `infer` call for the whole machine (src/ProbabilisticGraphicalModels/inference.jl:48-49).

The fan-out over GPUs and the one collective the sampling paths need live INSIDE libbn_cuda.so (csrc/comm.cu: NCCL
called directly from C).  This module is the thin host mirror of that part of the C ABI -- no torch, no
torch.distributed in the data path:

    comm = Communicator.init_all([0, 1, ..., 7])      # ONE process drives every GPU (what a Julia session does)
    comm = Communicator.from_env()                    # one process PER GPU (torchrun / mpirun): RANK, WORLD_SIZE,
                                                      #   LOCAL_RANK from the environment
    set_default_comm(comm)                            # every infer() now uses all GPUs of the communicator
    infer(LikelihoodWeightingInference(10**11), bn, query, evidence=e)

Sharding contract (include/bn_cuda.h): GPU g of G handles the global indices
[first + n*g//G, first + n*(g+1)//G).  Because the Philox counter is the global particle / chain index, the all-reduced
tensor equals the single-GPU tensor up to fp64 addition order (exactly, for the integer Gibbs counts).

`shard_range`, `allreduce_sum_` and `allgather_rows_` at the bottom are what the world-size-2 gloo tests on CPU use to
check that contract without a GPU; they are not on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile
import time
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import DeviceBuffer, check, ptr
from .device_layout import device_net


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first index, count) of rank's contiguous shard; shards tile [0, n_total) exactly.  Same rule as bn_comm_shard
    (csrc/comm.cu), restated in Python so that it can be checked without the library."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi - lo


def comm_shard(first: int, n: int, rank: int, world: int) -> Tuple[int, int]:
    """The C library's own sharding rule (bn_comm_shard; host-only, no GPU needed)."""
    lo, cnt = C.c_uint64(0), C.c_uint64(0)
    check(_lib.lib().bn_comm_shard(C.c_uint64(first), C.c_uint64(n), C.c_int32(rank), C.c_int32(world),
                                   C.byref(lo), C.byref(cnt)))
    return int(lo.value), int(cnt.value)


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


# ---------------------------------------------------------------------------------------------- unique-id exchange
_comm_seq = [0]


def _file_exchange(rank: int, world: int, make_id: Callable[[], bytes], timeout: float = 300.0) -> bytes:
    """Ship rank 0's 128-byte NCCL id to the other ranks of ONE node through a file (stdlib only): rank 0 writes
    `<dir>/bn_comm_<job>_<seq>.id` atomically, the others poll for it.  <job> = MASTER_ADDR, MASTER_PORT and the parent
    pid (torchrun's agent is the parent of every worker), <seq> = how many communicators this process has made.  A
    multi-node launch passes its own `exchange` (MPI bcast, a TCP store) to Communicator.from_env instead."""
    _comm_seq[0] += 1
    d = os.environ.get("BN_COMM_DIR", tempfile.gettempdir())
    job = os.environ.get("BN_COMM_JOB") or "{}_{}_{}".format(os.environ.get("MASTER_ADDR", "local"),
                                                             os.environ.get("MASTER_PORT", "0"), os.getppid())
    path = os.path.join(d, f"bn_comm_{job}_{_comm_seq[0]}.id")
    if rank == 0:
        uid = make_id()
        tmp = f"{path}.tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                uid = f.read()
            if len(uid) == _lib.BN_COMM_ID_BYTES:
                return uid
        except FileNotFoundError:
            pass
        if time.time() - t0 > timeout:
            raise TimeoutError(f"rank {rank}: no NCCL id from rank 0 at {path} after {timeout:.0f} s")
        time.sleep(0.005)


class Communicator:
    """bn_comm_t: the GPUs one sharded call runs on.  `world` GPUs in total, `n_local` of them driven by this process
    (global ranks first_rank .. first_rank + n_local - 1), `devices` = their CUDA ordinals."""

    def __init__(self, handle: int, devices: Sequence[int], cleanup: Optional[str] = None):
        import weakref
        self.handle = int(handle)
        self.devices = [int(d) for d in devices]
        w, f, n = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(_lib.lib().bn_comm_info(_lib.h64(self.handle), C.byref(w), C.byref(f), C.byref(n)))
        self.world, self.first_rank, self.n_local = int(w.value), int(f.value), int(n.value)
        self._fin = weakref.finalize(self, Communicator._destroy, self.handle)

    @staticmethod
    def _destroy(handle):
        try:
            _lib.lib().bn_comm_destroy(_lib.h64(handle))
        except Exception:
            pass

    def close(self):
        self._fin()

    # ---- construction
    @classmethod
    def init_all(cls, devices: Optional[Sequence[int]] = None) -> "Communicator":
        """One process, many GPUs (ncclCommInitAll): devices default to every visible GPU."""
        devices = list(range(_lib.device_count())) if devices is None else [int(d) for d in devices]
        arr = np.asarray(devices, np.int32)
        h = _lib.h64(0)
        check(_lib.lib().bn_comm_init_all(ptr(arr, _lib.i32p), C.c_int32(arr.size), C.byref(h)))
        return cls(h.value, devices)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * _lib.BN_COMM_ID_BYTES)()
        check(_lib.lib().bn_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def init_rank(cls, uid: bytes, rank: int, world: int, device: int) -> "Communicator":
        """One process per GPU (ncclCommInitRank): collective over all `world` processes."""
        if len(uid) != _lib.BN_COMM_ID_BYTES:
            raise ValueError(f"NCCL id must be {_lib.BN_COMM_ID_BYTES} bytes")
        buf = (C.c_uint8 * _lib.BN_COMM_ID_BYTES).from_buffer_copy(uid)
        h = _lib.h64(0)
        check(_lib.lib().bn_comm_init_rank(buf, C.c_int32(rank), C.c_int32(world), C.c_int32(device), C.byref(h)))
        return cls(h.value, [device])

    @classmethod
    def from_env(cls, exchange: Optional[Callable[[int, int, Callable[[], bytes]], bytes]] = None) -> "Communicator":
        """RANK / WORLD_SIZE / LOCAL_RANK (torchrun, mpirun wrappers) -> a process-per-GPU communicator.  `exchange(rank,
        world, make_id) -> id` ships rank 0's id to everybody; default: a file in the temp dir (single node)."""
        rank, world, local = env_rank_world()
        if world == 1:
            return cls.init_all([local])
        uid = (exchange or _file_exchange)(rank, world, cls.unique_id)
        return cls.init_rank(uid, rank, world, local)

    # ---- collectives on user data (the sampling entry points issue their own inside the C call)
    def allreduce_(self, arr: np.ndarray, op: str = "sum") -> np.ndarray:
        """In-place element-wise reduction of one float64 / int64 host array per PROCESS (timing maxima, bookkeeping)."""
        if arr.dtype not in (np.float64, np.int64) or not arr.flags.c_contiguous:
            raise TypeError("allreduce_ takes a contiguous float64 or int64 array")
        check(_lib.lib().bn_comm_allreduce(_lib.h64(self.handle), C.c_void_p(arr.ctypes.data), C.c_int64(arr.size),
                                           C.c_int32(_lib.DTYPE_F64 if arr.dtype == np.float64 else _lib.DTYPE_I64),
                                           C.c_int32(_lib.REDUCE_SUM if op == "sum" else _lib.REDUCE_MAX)))
        return arr

    def allreduce_dev_(self, dev_ptr: int, count: int, dtype=np.float64, op: str = "sum") -> None:
        """The same on a device buffer of the (single) local GPU, enqueued on the current stream."""
        check(_lib.lib().bn_comm_allreduce_dev(_lib.h64(self.handle), C.c_void_p(dev_ptr), C.c_int64(count),
                                               C.c_int32(_lib.DTYPE_F64 if np.dtype(dtype) == np.float64 else _lib.DTYPE_I64),
                                               C.c_int32(_lib.REDUCE_SUM if op == "sum" else _lib.REDUCE_MAX)))

    def barrier(self) -> None:
        check(_lib.lib().bn_comm_barrier(_lib.h64(self.handle)))

    def shard(self, first: int, n: int, local_index: int = 0) -> Tuple[int, int]:
        return comm_shard(first, n, self.first_rank + local_index, self.world)


# ---------------------------------------------------------------------------------------------- resident jobs
class _ShardedBase:
    def _setup(self, bn, query, evidence, comm, device):
        self.comm = comm
        rank, world, local = env_rank_world()
        self.rank = comm.first_rank if comm is not None else 0
        self.world = comm.world if comm is not None else 1
        self.net = device_net(bn, device, comm) if comm is not None else device_net(bn, local if device is None else device)
        self.device = self.net.device
        lay = self.net.layout
        self.ev = lay.evidence_vector(evidence)
        self.q = lay.query_indices(query)
        self.shape = [int(lay.card[i]) for i in self.q]
        self.ncell = int(np.prod(self.shape)) if self.shape else 1
        return lay


class ShardedLikelihoodWeighting(_ShardedBase):
    """LW inference of `n_total` particles over every GPU of `comm`, result resident in HBM.

    step() is ONE C call (bn_lw_infer_dev on a comm-bearing net): every GPU runs its share of the global particle
    range and the (cells + 2) doubles (weighted counts | sum w | sum w^2) are all-reduced by NCCL inside the call, on
    the stream the kernel ran on; normalisation happens after the reduction, exactly where the reference normalises
    (src/Inference/likelihood.jl:56).  Asynchronous: posterior() / counts() synchronise."""

    def __init__(self, bn, query, evidence, n_total: int, seed: int = 0, comm: Optional[Communicator] = None,
                 device: int | None = None):
        self._setup(bn, query, evidence, comm, device)
        self.n_total, self.first, self.seed = int(n_total), 0, int(seed)
        self.out = DeviceBuffer((self.ncell + 2) * 8, self.device, zero=True)

    def step(self, first=None, count=None):
        """Enqueue the whole job (all GPUs + all-reduce); the result is left in self.out on the first local GPU."""
        qp = self.q if self.q.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_lw_infer_dev(_lib.h64(self.net.handle), ptr(self.ev, _lib.i8p), ptr(qp, _lib.i32p),
                                         C.c_int32(self.q.size),
                                         C.c_uint64(self.first if first is None else first),
                                         C.c_uint64(self.n_total if count is None else count),
                                         C.c_uint64(self.seed), C.c_void_p(self.out.ptr)))
        return self.out

    launch = step

    def counts(self) -> np.ndarray:
        """[weighted counts (ncell) | sum w | sum w^2] of the last step, on the host (synchronises)."""
        return self.out.download(np.float64, self.ncell + 2)

    def posterior(self) -> np.ndarray:
        c = self.counts()[:self.ncell]
        with np.errstate(invalid="ignore", divide="ignore"):
            return (c / c.sum()).reshape(self.shape, order="F")


class ShardedGibbs(_ShardedBase):
    """Gibbs inference (Nodewise / Full) whose CHAINS are split over the GPUs of `comm` (BASELINE config C4: 65 536
    chains over 8 GPUs).  GPU g runs chains [g*C//G, (g+1)*C//G) of one global chain index space (the Philox counter
    is the global chain index), then the integer count tensors are all-reduced inside the C call: the result equals
    the single-GPU counts exactly."""

    def __init__(self, bn, query, evidence, n_chains: int, nsamples: int, burn_in: int, thin: int, full: bool = False,
                 seed: int = 0, comm: Optional[Communicator] = None, device: int | None = None):
        self._setup(bn, query, evidence, comm, device)
        self.n_chains, self.first = int(n_chains), 0
        self.nsamples, self.burn_in, self.thin, self.full, self.seed = int(nsamples), int(burn_in), int(thin), bool(full), int(seed)
        self.out = DeviceBuffer(self.ncell * 8, self.device, zero=True)

    def step(self):
        qp = self.q if self.q.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_gibbs_infer_dev(_lib.h64(self.net.handle), ptr(self.ev, _lib.i8p), ptr(qp, _lib.i32p),
                                            C.c_int32(self.q.size), C.c_uint64(self.first), C.c_uint64(self.n_chains),
                                            C.c_int64(self.nsamples), C.c_int64(self.burn_in), C.c_int64(self.thin),
                                            C.c_int32(1 if self.full else 0), C.c_uint64(self.seed),
                                            C.c_void_p(self.out.ptr)))
        return self.out

    launch = step

    def counts(self) -> np.ndarray:
        return self.out.download(np.float64, self.ncell)

    def posterior(self) -> np.ndarray:
        c = self.counts()
        with np.errstate(invalid="ignore", divide="ignore"):
            return (c / c.sum()).reshape(self.shape, order="F")


class ShardedTableStatistics:
    """The learn-from-samples loop with the table's ROWS split over the processes of `comm` (one GPU per process):
    rank g draws rows [g*R//G, (g+1)*R//G) of one global row index space (the Philox counter is the global row index,
    so the union of the shards IS the single-GPU table), counts its shard where it lies (bn_statistics_dev) and the
    int64 count tensors are all-reduced (bn_comm_allreduce_dev) -- integers, so every rank holds exactly the single-GPU
    counts.  The data log-likelihood is then the counts dotted with the log CPTs (bn_logpdf_counts_dev): the same bits
    on every rank and for every number of ranks.  No data-path collective besides that one all-reduce of cpt_len int64."""

    def __init__(self, bn, n_rows_total: int, seed: int = 0, comm: Optional[Communicator] = None, device: int | None = None):
        self.comm = comm
        rank, world, local = env_rank_world()
        self.rank = comm.first_rank if comm is not None else 0
        self.world = comm.world if comm is not None else 1
        if comm is not None and comm.n_local != 1:
            raise NotImplementedError("row-sharded tables use one GPU per process")
        self.device = comm.devices[0] if comm is not None else (local if device is None else device)
        self.net = device_net(bn, self.device)
        lay = self.net.layout
        self.n_rows_total = int(n_rows_total)
        self.first, self.count = shard_range(self.n_rows_total, self.rank, self.world)
        self.seed = int(seed)
        self.pitch = max(self.count, 1)
        self.table = DeviceBuffer(lay.n * self.pitch, self.device)  # node-major uint8, pitch = row count
        self.n_bins = int(lay.cpt_off[-1])
        self.counts = DeviceBuffer(self.n_bins * 8, self.device, zero=True)
        self.logl = DeviceBuffer(8, self.device, zero=True)

    def sample(self):
        """rand(bn, n): this rank's rows, left in HBM (node-major uint8, one column of `count` bytes per node)."""
        if self.count:
            check(_lib.lib().bn_sample_dev(_lib.h64(self.net.handle), None, C.c_int32(0), C.c_int32(1), C.c_uint64(self.first),
                                           C.c_uint64(self.count), C.c_uint64(self.seed), C.c_void_p(self.table.ptr),
                                           None, None))
        return self.table

    def launch(self):
        """Count the local shard only (no collective); int64 counts in self.counts."""
        check(_lib.lib().bn_statistics_dev(_lib.h64(self.net.handle), C.c_void_p(self.table.ptr), C.c_uint64(self.count),
                                           C.c_uint64(self.pitch), C.c_void_p(self.counts.ptr)))
        return self.counts

    def step(self):
        """sample -> count -> all-reduce: the whole table's sufficient statistics on every rank."""
        self.sample()
        self.launch()
        if self.comm is not None:
            self.comm.allreduce_dev_(self.counts.ptr, self.n_bins, np.int64)
        return self.counts

    def counts_host(self) -> np.ndarray:
        return self.counts.download(np.int64, self.n_bins)

    def logpdf(self) -> float:
        """logpdf(bn, table) of the WHOLE table from the all-reduced counts (call after step())."""
        check(_lib.lib().bn_logpdf_counts_dev(_lib.h64(self.net.handle), C.c_void_p(self.counts.ptr),
                                              C.c_void_p(self.logl.ptr)))
        return float(self.logl.download(np.float64, 1)[0])


class ShardedLoopy:
    """LoopyBelief over many (query, evidence) pairs, the PAIRS split over the processes of `comm`: rank g answers
    pairs [g*B//G, (g+1)*B//G) with one bn_loopy_infer_dev launch; pairs are independent, so there is no data-path
    collective (weak scaling)."""

    def __init__(self, bn, queries, evidences, nsamples: int = 500, tol: float = 1e-8, iters_for_convergence: int = 6,
                 comm: Optional[Communicator] = None, device: int | None = None):
        self.comm = comm
        rank, world, local = env_rank_world()
        self.rank = comm.first_rank if comm is not None else 0
        self.world = comm.world if comm is not None else 1
        self.device = comm.devices[0] if comm is not None else (local if device is None else device)
        self.net = device_net(bn, self.device)
        lay = self.net.layout
        self.n_total = len(queries)
        self.first, self.count = shard_range(self.n_total, self.rank, self.world)
        self.counts = [shard_range(self.n_total, g, self.world)[1] for g in range(self.world)]
        mine = range(self.first, self.first + self.count)
        ev = np.stack([lay.evidence_vector(evidences[b]) for b in mine]) if self.count else np.zeros((0, lay.n), np.int8)
        q = np.array([lay.index[queries[b]] for b in mine], np.int32)
        self.stride = int(lay.card.max())
        self.ev_d = DeviceBuffer(max(ev.nbytes, 1), self.device).upload(ev)
        self.q_d = DeviceBuffer(max(q.nbytes, 1), self.device).upload(q)
        self.out = DeviceBuffer(max(self.count * self.stride * 8, 8), self.device, zero=True)
        self.iters = DeviceBuffer(max(self.count * 4, 4), self.device, zero=True)
        self.nsamples, self.tol, self.K = int(nsamples), float(tol), int(iters_for_convergence)

    def launch(self):
        """Enqueue the local shard on the current stream; beliefs in self.out [count, max_card]."""
        check(_lib.lib().bn_loopy_infer_dev(_lib.h64(self.net.handle), C.c_void_p(self.ev_d.ptr),
                                            C.c_void_p(self.q_d.ptr), C.c_uint64(self.count), C.c_int32(self.nsamples),
                                            C.c_double(self.tol), C.c_int32(self.K), C.c_int32(self.stride),
                                            C.c_void_p(self.out.ptr), C.c_void_p(self.iters.ptr)))
        return self.out

    def beliefs(self) -> np.ndarray:
        return self.out.download(np.float64, self.count * self.stride).reshape(self.count, self.stride)


# ---------------------------------------------------------------------------------------------- CPU test helpers
def allreduce_sum_(t):
    """In-place sum over the ranks of a torch.distributed group (gloo on CPU): used by tests/test_distributed_cpu.py to
    check the sharding contract without a GPU.  The product path reduces inside libbn_cuda.so instead."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allgather_rows_(local, counts):
    """Concatenate per-rank row blocks over a torch.distributed group (CPU test helper, see allreduce_sum_)."""
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
