"""Factor: a dense named-dimension Float64 table living in HBM (mirror of src/Factors).

`*`, `sum`, `normalize!`, `ϕ[assignment]` and `Factor(bn, name, evidence)` are kernels behind the
C ABI (bn_factor_*).  `.potential` downloads a column-major numpy array shaped like the reference's
`Array{Float64,N}`.  States in assignments are 1-based, as in the reference.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import check, ptr, DimensionMismatch  # noqa: F401
from .device_layout import default_device, device_net

# Symbol <-> int32 id (the Julia glue keeps the same kind of table)
_ids: dict = {}
_names: list = []


def var_id(name) -> int:
    i = _ids.get(name)
    if i is None:
        i = len(_names)
        _ids[name] = i
        _names.append(name)
    return i


def var_name(i: int):
    return _names[i]


def _as_names(dims) -> list:
    if isinstance(dims, (str, bytes)) or not isinstance(dims, Iterable):
        return [dims]
    return list(dims)


_OPS = {"*": _lib.OP_MUL, "/": _lib.OP_DIV, "+": _lib.OP_ADD, "-": _lib.OP_SUB, "max": _lib.OP_MAX, "min": _lib.OP_MIN}


class Factor:
    """Factor(dims, potential) | Factor(dims, lengths, fill_value=0)   (src/Factors/factors_main.jl:18-56)."""

    def __init__(self, dims, potential_or_lengths=None, fill_value=0.0, *, device: int | None = None, _handle=None):
        self._fin = None
        if _handle is not None:
            self._adopt(_handle)
            return
        dims = _as_names(dims)
        if len(set(dims)) != len(dims):
            raise ValueError("Dimensions must be unique")  # errors.jl:10
        arg = potential_or_lengths
        if isinstance(arg, np.ndarray) and arg.dtype.kind == "f":
            pot = np.asarray(arg, np.float64)
        elif isinstance(arg, (list, tuple, np.ndarray)) and all(isinstance(x, (int, np.integer)) for x in arg):
            lengths = [int(x) for x in arg]
            if len(dims) != len(lengths):
                raise DimensionMismatch("`potential` must have as many dimensions as dims")
            pot = np.full(lengths, 0.0 if fill_value is None else float(fill_value), np.float64)
        else:
            pot = np.asarray(arg, np.float64)
        if len(dims) != pot.ndim:
            raise DimensionMismatch("`potential` must have as many dimensions as dims")  # factors_main.jl:27-29
        device = default_device() if device is None else device
        ids = np.array([var_id(d) for d in dims], np.int32)
        card = np.array(pot.shape, np.int32)
        flat = np.ascontiguousarray(pot.reshape(-1, order="F"))
        h = _lib.h64(0)
        dp = ids if ids.size else np.zeros(1, np.int32)
        cp = card if card.size else np.ones(1, np.int32)
        check(_lib.lib().bn_factor_upload(C.c_int32(len(dims)), ptr(dp, _lib.i32p), ptr(cp, _lib.i32p),
                                          ptr(flat, _lib.f64p), C.c_int32(device), C.byref(h)))
        self._adopt(h.value)

    # -- handle plumbing ---------------------------------------------------------------------
    def _adopt(self, handle: int) -> None:
        self.handle = int(handle)
        nd = C.c_int32(0)
        check(_lib.lib().bn_factor_ndims(_lib.h64(self.handle), C.byref(nd)))
        ids = np.zeros(max(nd.value, 1), np.int32)
        card = np.zeros(max(nd.value, 1), np.int32)
        check(_lib.lib().bn_factor_shape(_lib.h64(self.handle), ptr(ids, _lib.i32p), ptr(card, _lib.i32p)))
        self.dimensions = [var_name(int(i)) for i in ids[:nd.value]]
        self._shape = tuple(int(c) for c in card[:nd.value])
        self._fin = weakref.finalize(self, Factor._free, self.handle)

    @staticmethod
    def _free(handle):
        try:
            _lib.lib().bn_factor_free(_lib.h64(handle))
        except Exception:
            pass

    @classmethod
    def _from_handle(cls, h) -> "Factor":
        return cls(None, _handle=h.value if hasattr(h, "value") else h)

    # -- construction from a network node: Factor(bn, name, evidence) (factors_main.jl:79-83) ---
    @classmethod
    def from_bn(cls, bn, name, evidence: dict | None = None, device: int | None = None) -> "Factor":
        net = device_net(bn, device)
        # dims of CPD factors are node indices on the device; give them this network's names
        h = _lib.h64(0)
        check(_lib.lib().bn_factor_from_cpd(_lib.h64(net.handle), C.c_int32(net.layout.index[name]), C.byref(h)))
        cpd = bn.get(name)
        ids = np.array([var_id(n) for n in [name] + list(cpd.parents)], np.int32)
        check(_lib.lib().bn_factor_relabel(h, ptr(ids, _lib.i32p)))
        f = Factor._from_handle(h)
        return f[evidence] if evidence else f

    # -- basic queries (factors_main.jl:93-135) -----------------------------------------------
    @property
    def potential(self) -> np.ndarray:
        n = int(np.prod(self._shape)) if self._shape else 1
        out = np.zeros(n, np.float64)
        check(_lib.lib().bn_factor_download(_lib.h64(self.handle), ptr(out, _lib.f64p)))
        return out.reshape(self._shape, order="F")

    def names(self):
        return list(self.dimensions)

    @property
    def ndims(self) -> int:
        return len(self.dimensions)

    def size(self, *dims):
        if not dims:
            return self._shape
        for d in dims:
            if d not in self.dimensions:
                raise ValueError(f"{d} is not a valid dimension")
        r = tuple(self._shape[self.dimensions.index(d)] for d in dims)
        return r[0] if len(r) == 1 else r

    def __len__(self):
        return int(np.prod(self._shape)) if self._shape else 1

    def __contains__(self, dim):
        return dim in self.dimensions

    def _check_dims_valid(self, dims):  # auxiliary.jl:7-14
        for d in dims:
            if d not in self.dimensions:
                raise ValueError(f"{d} is not a valid dimension")

    # -- `*` = join(*, phi1, phi2, :outer) (factors_dims.jl:185-234,269) -----------------------
    def __mul__(self, other: "Factor") -> "Factor":
        if not isinstance(other, Factor):
            return NotImplemented
        h = _lib.h64(0)
        check(_lib.lib().bn_factor_product(_lib.h64(self.handle), _lib.h64(other.handle), C.byref(h)))
        return Factor._from_handle(h)

    # -- / + - = join(op, phi1, phi2, :outer) (factors_dims.jl:185-234,270-272) --------------------
    def _join(self, other: "Factor", op: int) -> "Factor":
        """NB the reference makes the larger factor phi1 BEFORE applying op (:189-191), so phi1 / phi2 with a smaller
        phi1 is phi2 ./ phi1 there; reproduced as is."""
        if not isinstance(other, Factor):
            return NotImplemented
        h = _lib.h64(0)
        check(_lib.lib().bn_factor_join(_lib.h64(self.handle), _lib.h64(other.handle), C.c_int32(op), C.byref(h)))
        return Factor._from_handle(h)

    def __truediv__(self, other):
        return self._join(other, _lib.OP_DIV)

    def __add__(self, other):
        return self._join(other, _lib.OP_ADD)

    def __sub__(self, other):
        return self._join(other, _lib.OP_SUB)

    def join(self, op, other: "Factor", kind: str = "outer") -> "Factor":
        """join(op, phi1, phi2, kind) with op one of '*', '/', '+', '-' (inner joins are unimplemented in the
        reference too, factors_dims.jl:236)."""
        if kind == "inner":
            raise RuntimeError("Inner joins are (still) umimplemented")
        if kind != "outer":
            raise ValueError(f"{kind} is not a supported join type")
        return self._join(other, _OPS[op])

    # -- reducedim(op, phi, dims) for * max min (factors_dims.jl:60-85,102-107) --------------------
    def reducedim(self, op, dims) -> "Factor":
        dims = _as_names(dims)
        self._check_dims_valid(dims)
        ids = np.array([var_id(d) for d in dims], np.int32)
        h = _lib.h64(0)
        dp = ids if ids.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_factor_reduce(_lib.h64(self.handle), ptr(dp, _lib.i32p), C.c_int32(ids.size),
                                          C.c_int32(_OPS[op]), C.byref(h)))
        return Factor._from_handle(h)

    def prod(self, dims) -> "Factor":
        return self.reducedim("*", dims)

    def maximum(self, dims) -> "Factor":
        return self.reducedim("max", dims)

    def minimum(self, dims) -> "Factor":
        return self.reducedim("min", dims)

    # -- sum(phi, dims) = reducedim(+, ...) (factors_dims.jl:60-85,100) -------------------------
    def sum(self, dims) -> "Factor":
        dims = _as_names(dims)
        self._check_dims_valid(dims)
        ids = np.array([var_id(d) for d in dims], np.int32)
        h = _lib.h64(0)
        dp = ids if ids.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_factor_sum(_lib.h64(self.handle), ptr(dp, _lib.i32p), C.c_int32(ids.size), C.byref(h)))
        return Factor._from_handle(h)

    # -- normalize!(phi; p) / normalize(phi; p) (factors_dims.jl:24-57) -------------------------
    def normalize_(self, dims=None, p: int = 1) -> "Factor":
        if p not in (1, 2):
            raise ValueError(f"p = {p} is not supported")
        if dims is not None:  # normalize!(phi, dims; p): factors_dims.jl:24-43
            dims = list(dict.fromkeys(_as_names(dims)))
            self._check_dims_valid(dims)
            ids = np.array([var_id(d) for d in dims], np.int32)
            dp = ids if ids.size else np.zeros(1, np.int32)
            check(_lib.lib().bn_factor_normalize_dims(_lib.h64(self.handle), ptr(dp, _lib.i32p), C.c_int32(ids.size),
                                                      C.c_int32(p)))
            return self
        check(_lib.lib().bn_factor_normalize(_lib.h64(self.handle), C.c_int32(p)))
        return self

    def copy(self) -> "Factor":
        h = _lib.h64(0)
        check(_lib.lib().bn_factor_slice(_lib.h64(self.handle), None, None, C.c_int32(0), C.byref(h)))
        return Factor._from_handle(h)

    def normalize(self, dims=None, p: int = 1) -> "Factor":
        if dims is not None:
            self._check_dims_valid(list(dict.fromkeys(_as_names(dims))))
        return self.copy().normalize_(dims, p)

    # -- phi[assignment] (factors_access.jl:19-35,48-72) ----------------------------------------
    def __getitem__(self, a):
        if isinstance(a, dict):
            dims, states = [], []
            for d in self.dimensions:
                if d in a:
                    v = a[d]
                    if v is Ellipsis or (isinstance(v, slice) and v == slice(None)):
                        continue
                    if not isinstance(v, (int, np.integer)) or isinstance(v, bool):
                        raise TypeError(f"Invalid state for dimension {d}")
                    n = self._shape[self.dimensions.index(d)]
                    if v < 1 or v > n:
                        raise IndexError(f"BoundsError({d}, {v})")
                    dims.append(var_id(d))
                    states.append(int(v) - 1)
            di = np.array(dims if dims else [0], np.int32)
            si = np.array(states if states else [0], np.int32)
            h = _lib.h64(0)
            check(_lib.lib().bn_factor_slice(_lib.h64(self.handle), ptr(di, _lib.i32p), ptr(si, _lib.i32p),
                                             C.c_int32(len(dims)), C.byref(h)))
            return Factor._from_handle(h)
        if isinstance(a, (int, np.integer)):
            return float(self.potential.reshape(-1, order="F")[a - 1])
        return self.pattern(a)

    # -- phi[assignment] = v (setindex!, factors_access.jl:39-46) -----------------------------------
    def __setitem__(self, a, v) -> None:
        if not isinstance(a, dict):
            raise TypeError("Factor[...] = v takes an Assignment (dict)")
        dims, states, free = [], [], []
        for i, d in enumerate(self.dimensions):
            if d in a and not (a[d] is Ellipsis or (isinstance(a[d], slice) and a[d] == slice(None))):
                x = a[d]
                if not isinstance(x, (int, np.integer)) or isinstance(x, bool):
                    raise TypeError(f"Invalid state for dimension {d}")
                if x < 1 or x > self._shape[i]:
                    raise IndexError(f"BoundsError({d}, {x})")
                dims.append(var_id(d))
                states.append(int(x) - 1)
            else:
                free.append(self._shape[i])
        vals = np.asarray(v, np.float64)
        if vals.ndim > 0:  # `.= v`: an array broadcasts over the slab; accepted here: a scalar-like or the slab's shape
            if vals.size != 1 and tuple(x for x in vals.shape if x != 1) != tuple(x for x in free if x != 1):
                raise DimensionMismatch(f"array could not be broadcast to match destination: {vals.shape} into {tuple(free)}")
        flat = np.ascontiguousarray(vals.reshape(-1, order="F"))
        di = np.array(dims if dims else [0], np.int32)
        si = np.array(states if states else [0], np.int32)
        check(_lib.lib().bn_factor_setindex(_lib.h64(self.handle), ptr(di, _lib.i32p), ptr(si, _lib.i32p),
                                            C.c_int32(len(dims)), ptr(flat, _lib.f64p), C.c_int64(flat.size)))

    # -- in-place variants rebind the device table, as the reference rebinds phi.potential ----------
    def _replace(self, other: "Factor") -> "Factor":
        if self._fin is not None:
            self._fin()  # frees the old table
        h = other.handle
        other._fin.detach()
        self._adopt(h)
        return self

    def permutedims(self, perm) -> "Factor":
        """permutedims(phi, perm) (factors_main.jl:160-166); perm holds 1-based positions like Julia's."""
        perm = [int(x) for x in perm]
        if sorted(perm) != list(range(1, len(self.dimensions) + 1)):
            raise ValueError("no valid permutation of dimensions")  # Base.permutedims: ArgumentError
        pi = np.array([x - 1 for x in perm] if perm else [0], np.int32)
        h = _lib.h64(0)
        check(_lib.lib().bn_factor_permutedims(_lib.h64(self.handle), ptr(pi, _lib.i32p), C.c_int32(len(perm)), C.byref(h)))
        return Factor._from_handle(h)

    def permutedims_(self, perm) -> "Factor":
        return self._replace(self.permutedims(perm))

    def push_(self, dim, length: int) -> "Factor":
        """push!(phi, dim, length) (factors_main.jl:148-158): appends a dimension along which the table repeats."""
        if dim in self.dimensions:
            raise RuntimeError(f"Dimension {dim} already exists")  # error(...), :150
        h = _lib.h64(0)
        check(_lib.lib().bn_factor_push(_lib.h64(self.handle), C.c_int32(var_id(dim)), C.c_int32(int(length)), C.byref(h)))
        return self._replace(Factor._from_handle(h))

    def sum_(self, dims) -> "Factor":      # sum!, prod!, maximum!, minimum! (factors_dims.jl:87-107)
        return self._replace(self.sum(dims))

    def prod_(self, dims) -> "Factor":
        return self._replace(self.prod(dims))

    def maximum_(self, dims) -> "Factor":
        return self._replace(self.maximum(dims))

    def minimum_(self, dims) -> "Factor":
        return self._replace(self.minimum(dims))

    def rand_(self, seed=None) -> "Factor":
        """rand!(phi) (factors_main.jl:140-143): fill with uniform [0, 1) values.  The reference draws from Julia's global
        RNG (a stream no reference test pins); here the values come from numpy and go down as one full-slab setindex!."""
        self[{}] = np.random.default_rng(seed).random(self._shape if self._shape else ())
        return self

    def similar(self) -> "Factor":
        """similar(phi) (factors_main.jl:93): same dims and sizes, contents unspecified (zeros here)."""
        return Factor(self.dimensions, list(self._shape))

    def indexin(self, dims):
        """indexin(dim(s), phi) (factors_main.jl:134-135): 1-based positions, None where absent."""
        if isinstance(dims, (str, bytes)) or not isinstance(dims, Iterable):
            return self.dimensions.index(dims) + 1 if dims in self.dimensions else None
        return [self.dimensions.index(d) + 1 if d in self.dimensions else None for d in dims]

    # -- broadcast(*, phi, dims, values) (factors_dims.jl:130-168): a product with 1-d factors ----
    def broadcast_mul(self, dims, values) -> "Factor":
        if isinstance(dims, (str, bytes)) or not isinstance(dims, Iterable):
            dims, values = [dims], [values]
        dims = list(dims)
        if len(set(dims)) != len(dims):
            raise ValueError("Dimensions must be unique")
        self._check_dims_valid(dims)
        if len(dims) != len(values):
            raise ValueError("Number of dimensions does not match number of values to broadcast")
        out = self
        for d, v in zip(dims, values):
            n = self._shape[self.dimensions.index(d)]
            if np.isscalar(v):
                vec = np.full(n, float(v))
            else:
                vec = np.asarray(v, np.float64)
                if vec.shape != (n,):
                    raise DimensionMismatch(f"broadcast value for {d} has length {vec.size}, dimension has {n}")
            out = out * Factor([d], vec)
        return out if out is not self else self.copy()

    def broadcast_(self, op, dims, values) -> "Factor":
        """broadcast!(op, phi, dims, values) in place (factors_dims.jl:131-165); op one of '*', '/', '+', '-', 'max', 'min'."""
        if isinstance(dims, (str, bytes)) or not isinstance(dims, Iterable):
            dims, values = [dims], [values]
        dims = list(dims)
        if len(set(dims)) != len(dims):
            raise ValueError("Dimensions must be unique")
        self._check_dims_valid(dims)
        if len(dims) != len(values):
            raise ValueError("Number of dimensions does not match number of values to broadcast")
        vecs = [np.atleast_1d(np.asarray(v, np.float64)) for v in values]
        for d, v in zip(dims, vecs):
            n = self._shape[self.dimensions.index(d)]
            if v.ndim != 1 or v.size not in (1, n):
                raise DimensionMismatch(f"broadcast value for {d} has length {v.size}, dimension has {n}")
        if not dims:
            return self
        ids = np.array([var_id(d) for d in dims], np.int32)
        lens = np.array([v.size for v in vecs], np.int32)
        flat = np.ascontiguousarray(np.concatenate(vecs))
        check(_lib.lib().bn_factor_broadcast(_lib.h64(self.handle), ptr(ids, _lib.i32p), C.c_int32(ids.size),
                                             ptr(flat, _lib.f64p), ptr(lens, _lib.i32p), C.c_int32(_OPS[op])))
        return self

    def broadcast(self, op, dims, values) -> "Factor":
        """broadcast(op, phi, dims, values) = broadcast!(op, deepcopy(phi), ...) (factors_dims.jl:118-119)."""
        return self.copy().broadcast_(op, dims, values)

    # -- pattern (factors_main.jl:174-199): host-side index bookkeeping ---------------------------
    def pattern(self, dims=None) -> np.ndarray:
        dims = self.dimensions if dims is None else _as_names(dims)
        self._check_dims_valid(dims)
        lens = list(self._shape)
        n = len(self)
        cols = []
        for d in dims:
            i = self.dimensions.index(d)
            inner = int(np.prod(lens[:i])) if i else 1
            outer = n // lens[i] // inner
            cols.append(np.tile(np.repeat(np.arange(1, lens[i] + 1), inner), outer))
        return np.stack(cols, axis=1)

    def to_dataframe(self):
        """convert(DataFrame, phi) (factors_dataframes.jl:9-20)."""
        import pandas as pd
        pat = self.pattern()
        df = pd.DataFrame({d: pat[:, i] for i, d in enumerate(self.dimensions)})
        df["potential"] = self.potential.reshape(-1, order="F")
        return df

    def __repr__(self):
        return f"Factor(dims={self.dimensions}, size={self._shape})"


def eliminate(factors: Sequence[Factor], dims) -> Factor:
    """sum(reduce(*, factors), dims) in ONE kernel: the elimination step of src/Inference/exact.jl:27 without
    materialising the product.  Values and dimension order equal the reference's left fold followed by sum."""
    factors = list(factors)
    if not factors:
        raise ValueError("eliminate needs at least one factor")
    ids = np.array([var_id(d) for d in _as_names(dims)], np.int32)
    hs = (_lib.h64 * len(factors))(*[f.handle for f in factors])
    out = _lib.h64(0)
    dp = ids if ids.size else np.zeros(1, np.int32)
    check(_lib.lib().bn_factor_eliminate(hs, C.c_int32(len(factors)), ptr(dp, _lib.i32p), C.c_int32(ids.size),
                                         C.byref(out)))
    return Factor._from_handle(out)
