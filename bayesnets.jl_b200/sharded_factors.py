"""A Factor split over the GPUs of one box along its OUTERMOST variables (the optional part of
north_star: "Factor ops stay single-GPU unless the largest factor is split along an outer variable").

One process per GPU (torch.distributed; NCCL over NVLink / NVSwitch on the box, gloo in the CPU tests).  Rank g holds
the slice of the factor where the split variables take their g-th joint state (first split variable fastest), as an
ordinary single-GPU Factor without those dims.  With that layout

  * `*` with a replicated factor, or with a factor split the same way, is LOCAL (a replicated operand that mentions a
    split variable is first sliced at this rank's state -- bn_factor_slice);
  * `sum` over variables that are not split is LOCAL;
  * `sum` over the split variables is the one real exchange step of the whole inference path: the ranks' tables are the
    addends, so it is ONE all-reduce(SUM) of the local table (2^k doubles) and the result is replicated;
  * `normalize!` is the local sum(abs) (bn_factor_norm), a scalar all-reduce, and the local `./=` (bn_factor_div_scalar);
  * the fused `sum(reduce(*, fs), h)` elimination step is local for any h that is not split.

Values equal the single-GPU factor's: products and local sums bit for bit; the all-reduce adds the ranks' addends in
NCCL's order instead of the reference's first-state-first order, so a sum over a split variable with more than two
states can differ in the last bit (two states: exact, fp addition commutes) -- inside the 1e-12 bar either way.

The class is written against the small factor interface (`dimensions`, `*`, `sum`, slicing by assignment, `potential`)
shared by bayesnets.jl_b200.Factor and the numpy oracle, so the world-size-2 gloo test drives the same sharding logic on
CPU with the oracle's factors as the backend.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


class FactorBackend:
    """What ShardedFactor needs from a single-device factor implementation."""

    def slice(self, f, assignment: dict):           # {dim: 0-based state} -> factor without those dims
        raise NotImplementedError

    def mul(self, a, b):
        return a * b

    def sum(self, f, dims):
        return f.sum(list(dims))

    def eliminate(self, fs, dims):
        acc = fs[0]
        for x in fs[1:]:
            acc = acc * x
        return acc.sum(list(dims)) if dims else acc

    def allreduce_sum(self, f):                     # in place over the default process group; returns f
        raise NotImplementedError

    def norm(self, f, p: int) -> float:
        raise NotImplementedError

    def div_scalar(self, f, total: float):
        raise NotImplementedError

    def host(self, f) -> np.ndarray:
        return np.asarray(f.potential)


class DeviceBackend(FactorBackend):
    """bayesnets.jl_b200.Factor on this rank's GPU; collectives on the factor's own device memory."""

    def __init__(self):
        import torch
        from . import _lib
        from .factors import Factor, eliminate
        self.torch, self._lib, self.Factor, self._eliminate = torch, _lib, Factor, eliminate

    def slice(self, f, assignment):
        return f[{d: int(s) + 1 for d, s in assignment.items()}]  # the mirror takes 1-based states

    def eliminate(self, fs, dims):
        return self._eliminate(list(fs), list(dims))

    def _tensor(self, f):
        """Zero-copy torch view of the factor's device memory (for the collective only)."""
        import ctypes as C
        p = C.c_void_p(0)
        self._lib.check(self._lib.lib().bn_factor_dev_ptr(self._lib.h64(f.handle), C.byref(p)))
        n = int(np.prod(f._shape)) if f._shape else 1

        class _Iface:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(p.value), False), "version": 3}
        return self.torch.as_tensor(_Iface(), device=f"cuda:{self.torch.cuda.current_device()}")

    def allreduce_sum(self, f):
        dist = _dist()
        self.torch.cuda.current_stream().synchronize()  # the factor was produced on the library's stream
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._tensor(f), op=dist.ReduceOp.SUM)
            self.torch.cuda.synchronize()
        return f

    def norm(self, f, p):
        import ctypes as C
        out = C.c_double(0.0)
        self._lib.check(self._lib.lib().bn_factor_norm(self._lib.h64(f.handle), C.c_int32(p), C.byref(out)))
        return float(out.value)

    def div_scalar(self, f, total):
        import ctypes as C
        self._lib.check(self._lib.lib().bn_factor_div_scalar(self._lib.h64(f.handle), C.c_double(total)))
        return f


class ShardedFactor:
    """`local` = this rank's slice (split dims removed); `split` = [(dim, card), ...] with prod(card) == world size."""

    def __init__(self, local, split: Sequence[tuple], rank: int, world: int, backend: FactorBackend):
        cards = [int(c) for _, c in split]
        if int(np.prod(cards)) != world:
            raise ValueError(f"split cardinalities {cards} do not multiply to the world size {world}")
        for d, _ in split:
            if d in local.dimensions:
                raise ValueError(f"split variable {d} must not be a dimension of the local slice")
        self.local, self.split, self.rank, self.world, self.be = local, [(d, int(c)) for d, c in split], rank, world, backend

    # ---- construction ------------------------------------------------------------------------
    @staticmethod
    def rank_state(split, rank: int) -> dict:
        """0-based states of the split variables on `rank` (first split variable fastest)."""
        out, r = {}, rank
        for d, c in split:
            out[d] = r % c
            r //= c
        return out

    @classmethod
    def from_replicated(cls, full, split_dims: Sequence, rank: int, world: int, backend: FactorBackend) -> "ShardedFactor":
        """Every rank holds `full`; keep only this rank's slice."""
        shape = dict(zip(full.dimensions, np.asarray(full.potential).shape)) if not hasattr(full, "_shape") else \
            dict(zip(full.dimensions, full._shape))
        split = [(d, int(shape[d])) for d in split_dims]
        return cls(backend.slice(full, cls.rank_state(split, rank)), split, rank, world, backend)

    @property
    def dimensions(self):
        return list(self.local.dimensions) + [d for d, _ in self.split]

    def _mine(self) -> dict:
        return self.rank_state(self.split, self.rank)

    def _localise(self, other):
        """A replicated operand restricted to this rank's states of the split variables it mentions."""
        mine = {d: s for d, s in self._mine().items() if d in other.dimensions}
        return self.be.slice(other, mine) if mine else other

    # ---- algebra -----------------------------------------------------------------------------
    def __mul__(self, other) -> "ShardedFactor":
        if isinstance(other, ShardedFactor):
            if other.split != self.split:
                raise ValueError("factors are split differently")
            rhs = other.local
        else:
            rhs = self._localise(other)
        return ShardedFactor(self.be.mul(self.local, rhs), self.split, self.rank, self.world, self.be)

    __rmul__ = __mul__

    def sum(self, dims):
        """Local for unsplit variables; summing split variables needs ALL of them and returns a replicated factor."""
        dims = [dims] if not isinstance(dims, (list, tuple)) else list(dims)
        sdims = [d for d, _ in self.split]
        hit = [d for d in dims if d in sdims]
        local_dims = [d for d in dims if d not in sdims]
        for d in local_dims:
            if d not in self.local.dimensions:
                raise ValueError(f"{d} is not a valid dimension")
        out = self.be.sum(self.local, local_dims) if local_dims else self.local
        if not hit:
            return ShardedFactor(out, self.split, self.rank, self.world, self.be)
        if sorted(hit, key=str) != sorted(sdims, key=str):
            raise NotImplementedError("summing out some but not all split variables (would need sub-group collectives)")
        if out is self.local:
            out = self.be.sum(self.local, [])  # a private copy: the all-reduce works in place
        return self.be.allreduce_sum(out)

    def eliminate(self, others: Sequence, dims):
        """sum(reduce(*, [self, others...]), dims) in one pass per rank, for variables that are not split."""
        dims = [dims] if not isinstance(dims, (list, tuple)) else list(dims)
        if any(d in [s for s, _ in self.split] for d in dims):
            raise NotImplementedError("eliminate over a split variable: multiply, then sum")
        fs = [self.local] + [o.local if isinstance(o, ShardedFactor) else self._localise(o) for o in others]
        return ShardedFactor(self.be.eliminate(fs, dims), self.split, self.rank, self.world, self.be)

    def normalize_(self, p: int = 1) -> "ShardedFactor":
        import torch
        dist = _dist()
        t = torch.tensor([self.be.norm(self.local, p)], dtype=torch.float64)
        if dist.is_initialized() and dist.get_world_size() > 1:
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        self.be.div_scalar(self.local, float(t.item()))
        return self

    def gather(self) -> np.ndarray:
        """The full potential (host, dims = local dims then split dims) on every rank -- for checks, not for speed."""
        import torch
        dist = _dist()
        mine = torch.from_numpy(np.ascontiguousarray(self.be.host(self.local).reshape(-1, order="F")))
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        if dist.is_initialized() and dist.get_world_size() > 1:
            if dist.get_backend() == "nccl":
                mine = mine.cuda()
                parts = [p.cuda() for p in parts]
            dist.all_gather(parts, mine)
        else:
            parts = [mine]
        shape = tuple(np.asarray(self.be.host(self.local)).shape) + tuple(c for _, c in self.split)
        return torch.stack([p.cpu() for p in parts], dim=-1).numpy().reshape(shape, order="F")
