"""Flatten a DiscreteBayesNet into the compact device layout and upload it (bn_net_create).

Python twin of julia/device_layout.jl (the file a BayesNets.jl maintainer adds as
src/DiscreteBayesNet/device_layout.jl).  Layout (include/bn_cuda.h):
  nodes in bn.cpds order = topological (src/bayes_nets.jl:26-48)
  card[n]              ncategories per node                (discrete_bayes_net.jl:96-98)
  par_ptr/par_idx      CSR of parents in cpd.parents order (first parent fastest in the CPT)
  cpt_off[n+1], cpt    vcat(d.p for d in cpd.distributions) per node, node state fastest
                       (src/Factors/factors_main.jl:62-67)
States are shifted to 0-based here; evidence becomes an int8 vector with -1 = free.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, ptr


@dataclass
class DeviceLayout:
    names: list
    index: dict
    card: np.ndarray      # int32 [n]
    par_ptr: np.ndarray   # int32 [n+1]
    par_idx: np.ndarray   # int32 [sum parents]
    cpt_off: np.ndarray   # int64 [n+1]
    cpt: np.ndarray       # float64

    @property
    def n(self) -> int:
        return len(self.names)

    def parents(self, i: int):
        return [int(p) for p in self.par_idx[self.par_ptr[i]:self.par_ptr[i + 1]]]

    def evidence_vector(self, evidence: dict | None) -> np.ndarray:
        """Assignment (1-based states) -> int8[n], -1 = free.  Names not in the network are ignored,
        as `haskey(evidence, s) for s in nodes` does (src/Inference/likelihood.jl:25)."""
        ev = np.full(self.n, -1, np.int8)
        for k, v in (evidence or {}).items():
            if k not in self.index:
                continue
            i = self.index[k]
            if not isinstance(v, (int, np.integer)):
                raise TypeError(f"Invalid state for {k}: {v!r}")
            if not 1 <= int(v) <= int(self.card[i]):
                raise IndexError(f"BoundsError: state {v} of {k} outside 1:{int(self.card[i])}")
            if int(v) - 1 > _lib.BN_MAX_EVIDENCE_STATE:  # int8 on the ABI: never let it wrap to "free"
                raise IndexError(f"BoundsError: evidence state {v} of {k}: only states 1:{_lib.BN_MAX_EVIDENCE_STATE + 1} "
                                 "of a node can be clamped as evidence (int8 in include/bn_cuda.h)")
            ev[i] = int(v) - 1
        return ev

    def query_indices(self, query) -> np.ndarray:
        return np.array([self.index[q] for q in query], np.int32)


def flatten(bn) -> DeviceLayout:
    names = bn.names()
    index = {n: i for i, n in enumerate(names)}
    card, par_ptr, par_idx, cpt_off, chunks = [], [0], [], [0], []
    for i, cpd in enumerate(bn.cpds):
        k = cpd.ncategories()
        card.append(k)
        for p, pc in zip(cpd.parents, cpd.parental_ncategories):
            j = index[p]
            if j >= i:
                raise ValueError("BayesNet is not topologically ordered")
            if bn.cpds[j].ncategories() != pc:
                raise ValueError(f"CPD {cpd.name}: parent {p} has {bn.cpds[j].ncategories()} categories, CPD says {pc}")
            par_idx.append(j)
        par_ptr.append(len(par_idx))
        tab = np.concatenate([np.asarray(d, np.float64) for d in cpd.distributions])
        chunks.append(tab)
        cpt_off.append(cpt_off[-1] + tab.size)
    return DeviceLayout(names, index, np.asarray(card, np.int32), np.asarray(par_ptr, np.int32),
                        np.asarray(par_idx, np.int32), np.asarray(cpt_off, np.int64),
                        np.concatenate(chunks) if chunks else np.zeros(0))


class DeviceNet:
    """A DiscreteBayesNet resident in HBM (opaque bn_net_t handle).  With a Communicator (distributed.py) the network
    is replicated on every local GPU of the communicator (bn_net_create_comm) and the sampling entry points shard
    their particles / chains over all of its GPUs inside the C call; `device` is then the first local GPU."""

    def __init__(self, layout: DeviceLayout, device: int = 0, comm=None):
        if int(layout.card.max(initial=1)) > _lib.BN_MAX_CARD:
            raise ValueError(f"a node has {int(layout.card.max())} states; the device layout carries states as uint8 "
                             f"with 0xFF reserved, i.e. at most {_lib.BN_MAX_CARD} states per node")
        self.layout = layout
        self.comm = comm
        self.device = device if comm is None else comm.devices[0]
        h = _lib.h64(0)
        par_idx = layout.par_idx if layout.par_idx.size else np.zeros(1, np.int32)
        if comm is None:
            check(_lib.lib().bn_net_create(C.c_int32(layout.n), ptr(layout.card, _lib.i32p), ptr(layout.par_ptr, _lib.i32p),
                                           ptr(par_idx, _lib.i32p), ptr(layout.cpt_off, _lib.i64p),
                                           ptr(layout.cpt, _lib.f64p), C.c_int32(device), C.byref(h)))
        else:
            check(_lib.lib().bn_net_create_comm(C.c_int32(layout.n), ptr(layout.card, _lib.i32p),
                                                ptr(layout.par_ptr, _lib.i32p), ptr(par_idx, _lib.i32p),
                                                ptr(layout.cpt_off, _lib.i64p), ptr(layout.cpt, _lib.f64p),
                                                _lib.h64(comm.handle), C.byref(h)))
        self.handle = h.value
        self._fin = weakref.finalize(self, DeviceNet._destroy, self.handle)

    @staticmethod
    def _destroy(handle):
        try:
            _lib.lib().bn_net_destroy(_lib.h64(handle))
        except Exception:
            pass

    def close(self):
        self._fin()


_default_device = 0
_default_comm = None


def set_default_device(device: int) -> None:
    global _default_device
    _default_device = int(device)


def default_device() -> int:
    return _default_device


def set_default_comm(comm) -> None:
    """Make `comm` (a distributed.Communicator, or None) the default for device_net(): every infer() / raw_counts()
    that does not name a device then runs on all GPUs of the communicator -- one call, whole machine."""
    global _default_comm
    _default_comm = comm


def default_comm():
    return _default_comm


def _fingerprint(lay: DeviceLayout) -> bytes:
    """Content hash of what the device holds: structure + CPT bytes.  Mutating a CPD in place (or replacing one)
    changes it, so the stale device copy is never reused; ~10 us for a 100-node net."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for a in (lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update("\0".join(lay.names).encode())
    return h.digest()


def device_net(bn, device: int | None = None, comm=None) -> DeviceNet:
    """The network's device copy: uploaded once per (network CONTENTS, device or communicator).  The network is
    re-flattened on every call (host work, tens of microseconds) and the copy is keyed by a content fingerprint, so
    in-place edits of a CPD's distributions are seen; a few copies (e.g. one per device) are kept per network."""
    if comm is None and device is None:
        comm = _default_comm
    device = _default_device if device is None else device
    lay = flatten(bn)
    key = (_fingerprint(lay), ("comm", id(comm)) if comm is not None else ("dev", device))
    cache = getattr(bn, "_device_cache", None)
    if not isinstance(cache, dict):
        cache = {}
        bn._device_cache = cache
    net = cache.get(key)
    if net is None:
        for k in [k for k in cache if k[0] != key[0]]:  # contents changed: drop the stale copies
            cache.pop(k).close()
        net = DeviceNet(lay, device, comm)
        cache[key] = net
    return net
