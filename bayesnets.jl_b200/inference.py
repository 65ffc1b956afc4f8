"""infer(im, bn, query; evidence) for the offloaded InferenceMethods.

Mirrors src/ProbabilisticGraphicalModels/inference.jl:4-49 (InferenceState, generic infer entry) and the
`infer` methods of src/Inference/{likelihood,gibbs,exact,loopy_belief}.jl.  Each method flattens nothing itself: the
network is uploaded once (device_layout.device_net) and the inner loop is one C-ABI call.

Extra fields the reference does not have: `seed` (counter-based Philox key; the reference uses Julia's
global RNG), `n_chains` for Gibbs (1 = the reference's single chain) and `device`.

Seeds.  The reference draws fresh random numbers on every call.  Here `seed=None` (the default) takes the next value of
a process-wide counter per call; an explicit seed is used as given by the FIRST call made with it and advanced for
every further call on the same method object (so repeated `infer` calls, and Gibbs chains resumed through `im.state`,
never replay a stream); assigning `im.seed` again restarts from that value.  `last_seed` is what the last call used.
With set_default_comm(comm) (distributed.py) the same calls run on every GPU of the communicator.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check, ptr
from .device_layout import device_net
from .factors import Factor, _as_names


class InferenceMethod:
    """abstract type InferenceMethod (src/ProbabilisticGraphicalModels/inference.jl:4)."""


class InferenceState:
    """InferenceState(pgm, query, evidence): validates query ⊆ nodes, query ∩ evidence = ∅, and
    de-duplicates the query (src/ProbabilisticGraphicalModels/inference.jl:6-33)."""

    def __init__(self, pgm, query, evidence: Optional[dict] = None):
        evidence = dict(evidence or {})
        query = list(dict.fromkeys(_as_names(query)))  # unique(), order-preserving
        names = set(pgm.names())
        for q in query:
            if q not in names:
                raise ValueError(f"Query {q} is not in the probabilistic graphical model")
            if q in evidence:
                raise ValueError(f"Query {q} is part of the evidence")
        self.pgm, self.query, self.evidence = pgm, query, evidence

    def __repr__(self):
        return f"Query: {self.query}\nEvidence: {self.evidence}"


_auto_seed = [0]
_MASK64 = (1 << 64) - 1


class _SeedStream:
    """Per-call Philox key of a sampling InferenceMethod (module docstring, "Seeds")."""
    _seed_owner = None
    _seed_calls = 0
    last_seed: Optional[int] = None

    def _call_seed(self) -> int:
        if self.seed is None:
            _auto_seed[0] += 1
            s = 0x5EED000000000000 + _auto_seed[0]
        else:
            if self._seed_owner != self.seed:
                self._seed_owner, self._seed_calls = self.seed, 0
            k = self._seed_calls
            self._seed_calls += 1
            s = (int(self.seed) + k * 0x9E3779B97F4A7C15) & _MASK64
        self.last_seed = s
        return s


def _query_factor(inf: InferenceState, counts: np.ndarray, device) -> Factor:
    shape = [inf.pgm.ncategories(q) for q in inf.query]
    return Factor(inf.query, np.asarray(counts, np.float64).reshape(shape, order="F"), device=device)


# ---------------------------------------------------------------------------------------------
@dataclass
class LikelihoodWeightingInference(_SeedStream, InferenceMethod):
    """src/Inference/likelihood.jl:5-7 (+ seed/device).  `last_stats` = (sum w, sum w^2) of the last call."""
    nsamples: int = 500
    seed: Optional[int] = None
    device: Optional[int] = None
    first_particle: int = 0
    last_stats: tuple = field(default=(0.0, 0.0), repr=False, compare=False)

    def raw_counts(self, inf: InferenceState):
        """Unnormalised weighted counts of this call's particle range (what multi-GPU ranks all-reduce)."""
        net = device_net(inf.pgm, self.device)
        lay = net.layout
        ev = lay.evidence_vector(inf.evidence)
        q = lay.query_indices(inf.query)
        ncell = int(np.prod([lay.card[i] for i in q])) if len(q) else 1
        counts = np.zeros(ncell, np.float64)
        stats = np.zeros(2, np.float64)
        qp = q if q.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_lw_infer(_lib.h64(net.handle), ptr(ev, _lib.i8p), ptr(qp, _lib.i32p), C.c_int32(q.size),
                                     C.c_uint64(self.first_particle), C.c_uint64(self.nsamples), C.c_uint64(self._call_seed()),
                                     ptr(counts, _lib.f64p), ptr(stats, _lib.f64p)))
        self.last_stats = (float(stats[0]), float(stats[1]))
        return counts

    def infer(self, inf: InferenceState) -> Factor:
        """src/Inference/likelihood.jl:16-58: weighted counts, then normalize! (0/0 -> NaN, as the reference)."""
        net = device_net(inf.pgm, self.device)
        phi = _query_factor(inf, self.raw_counts(inf), net.device)
        return phi.normalize_()


@dataclass
class _GibbsBase(_SeedStream, InferenceMethod):
    nsamples: int = 2000
    burn_in: int = 500
    thin: int = 3
    state: dict = field(default_factory=dict)  # carried across calls, like im.state (gibbs.jl:64,83-88)
    n_chains: int = 1
    seed: Optional[int] = None
    device: Optional[int] = None
    first_chain: int = 0
    chain_states: Optional[np.ndarray] = field(default=None, repr=False, compare=False)
    last_chain_counts: Optional[np.ndarray] = field(default=None, repr=False, compare=False)
    keep_chain_counts: bool = False
    _mode = _lib.BN_OK  # overwritten

    def raw_counts(self, inf: InferenceState):
        if not self.burn_in < self.nsamples:  # gibbs.jl:70-71
            raise ValueError(f"Burn in ({self.burn_in}) must be less than number of samples ({self.nsamples})")
        net = device_net(inf.pgm, self.device)
        lay = net.layout
        ev = lay.evidence_vector(inf.evidence)
        q = lay.query_indices(inf.query)
        ncell = int(np.prod([lay.card[i] for i in q])) if len(q) else 1
        counts = np.zeros(ncell, np.float64)
        nch = int(self.n_chains)
        init = 0
        if self.chain_states is not None and self.chain_states.shape == (nch, lay.n):
            st = np.ascontiguousarray(self.chain_states, np.uint8)
            init = 1
        elif self.state:
            # a user-supplied / previous single-chain state (1-based Assignment), replicated over chains
            row = np.zeros(lay.n, np.uint8)
            for i, n in enumerate(lay.names):
                v = int(self.state[n])
                if not 1 <= v <= int(lay.card[i]):  # the state indexes the CPTs on the device
                    raise IndexError(f"BoundsError: state {v} of {n} outside 1:{int(lay.card[i])}")
                row[i] = v - 1
            for k, v in inf.evidence.items():
                if k in lay.index:
                    row[lay.index[k]] = int(v) - 1
            st = np.ascontiguousarray(np.tile(row, (nch, 1)))
            init = 1
        else:
            st = np.zeros((nch, lay.n), np.uint8)
        cc = np.zeros((nch, ncell), np.float64) if self.keep_chain_counts else None
        qp = q if q.size else np.zeros(1, np.int32)
        check(_lib.lib().bn_gibbs_infer(_lib.h64(net.handle), ptr(ev, _lib.i8p), ptr(qp, _lib.i32p), C.c_int32(q.size),
                                        C.c_uint64(self.first_chain), C.c_uint64(nch), C.c_int64(self.nsamples),
                                        C.c_int64(self.burn_in), C.c_int64(self.thin), C.c_int32(self._mode),
                                        C.c_uint64(self._call_seed()), ptr(st, _lib.u8p), C.c_int32(init),
                                        ptr(counts, _lib.f64p), ptr(cc, _lib.f64p)))
        self.chain_states = st
        self.state = {n: int(st[0, i]) + 1 for i, n in enumerate(lay.names)}
        self.last_chain_counts = cc
        return counts

    def infer(self, inf: InferenceState) -> Factor:
        counts = self.raw_counts(inf)  # validates burn_in < nsamples before touching the device
        net = device_net(inf.pgm, self.device)
        return _query_factor(inf, counts, net.device).normalize_()


@dataclass
class GibbsSamplingNodewise(_GibbsBase):
    """src/Inference/gibbs.jl:60-145: one node update = one iteration."""
    _mode = 0


@dataclass
class GibbsSamplingFull(_GibbsBase):
    """src/Inference/gibbs.jl:155-234: one sweep = one iteration."""
    _mode = 1


@dataclass
class ExactInference(InferenceMethod):
    """src/Inference/exact.jl:4-33.  fused=True sums each hidden variable out of the product of its
    factors in one kernel; fused=False reproduces the reference's pairwise products."""
    fused: bool = True
    device: Optional[int] = None

    def infer(self, inf: InferenceState) -> Factor:
        net = device_net(inf.pgm, self.device)
        lay = net.layout
        ev = lay.evidence_vector(inf.evidence)
        q = lay.query_indices(inf.query)
        if q.size == 0:
            raise ValueError("ExactInference needs at least one query variable")
        ncell = int(np.prod([lay.card[i] for i in q]))
        out = np.zeros(ncell, np.float64)
        od = np.zeros(q.size, np.int32)
        oc = np.zeros(q.size, np.int32)
        check(_lib.lib().bn_exact_infer(_lib.h64(net.handle), ptr(ev, _lib.i8p), ptr(q, _lib.i32p), C.c_int32(q.size),
                                        C.c_int32(1 if self.fused else 0), ptr(od, _lib.i32p), ptr(oc, _lib.i32p),
                                        ptr(out, _lib.f64p)))
        # the result's dim order is join-order dependent, as in the reference (compare by name)
        dims = [lay.names[int(i)] for i in od]
        return Factor(dims, out.reshape([int(c) for c in oc], order="F"), device=net.device)


@dataclass
class LoopyBelief(InferenceMethod):
    """src/Inference/loopy_belief.jl:7-11: nsamples = iteration cap, early stop once the largest message change
    stayed <= tol for iters_for_convergence iterations (tol < 0: never).  `last_iters` = iterations of the last call.

    Same numbers as the reference, including the effect of its message aliasing (csrc/loopy.cu header): exact on the
    reference's own test nets, not textbook LBP when an unobserved root has several children."""
    nsamples: int = 500
    tol: float = 1e-8
    iters_for_convergence: int = 6
    device: Optional[int] = None
    last_iters: Optional[np.ndarray] = field(default=None, repr=False, compare=False)

    def infer(self, inf: InferenceState) -> Factor:
        if len(inf.query) != 1:  # loopy_belief.jl:26
            raise ValueError("There can only be one query variable")
        post = self.infer_batch(inf.pgm, [inf.query[0]], [inf.evidence])
        net = device_net(inf.pgm, self.device)
        return Factor(inf.query, post[0, :inf.pgm.ncategories(inf.query[0])], device=net.device)

    def infer_batch(self, bn, queries, evidences) -> np.ndarray:
        """Many (query, evidence) pairs of one network in ONE kernel launch (one warp per pair).  queries: node names;
        evidences: Assignments (1-based states) or None.  Returns [n_pairs, max_card] posteriors, rows zero-padded."""
        net = device_net(bn, self.device)
        lay = net.layout
        queries = list(queries)
        evidences = list(evidences) if evidences is not None else [None] * len(queries)
        if len(evidences) != len(queries):
            raise ValueError("one evidence Assignment per query")
        infs = [InferenceState(bn, q, e) for q, e in zip(queries, evidences)]
        for i in infs:
            if len(i.query) != 1:
                raise ValueError("There can only be one query variable")
        B = len(infs)
        ev = np.empty((B, lay.n), np.int8)
        for b, i in enumerate(infs):
            ev[b] = lay.evidence_vector(i.evidence)
        q = np.array([lay.index[i.query[0]] for i in infs], np.int32)
        stride = int(lay.card.max())
        out = np.zeros((B, stride), np.float64)
        iters = np.zeros(B, np.int32)
        if B:
            check(_lib.lib().bn_loopy_infer(_lib.h64(net.handle), ptr(ev, _lib.i8p), ptr(q, _lib.i32p), C.c_uint64(B),
                                            C.c_int32(self.nsamples), C.c_double(self.tol),
                                            C.c_int32(self.iters_for_convergence), C.c_int32(stride),
                                            ptr(out, _lib.f64p), ptr(iters, _lib.i32p)))
        self.last_iters = iters
        return out


def infer(*args, evidence: Optional[dict] = None) -> Factor:
    """infer(im, inf) | infer(im, bn, query; evidence) | infer(bn, query; evidence) | infer(inf)
    (src/ProbabilisticGraphicalModels/inference.jl:48-49, src/Inference/exact.jl:34-35)."""
    if isinstance(args[0], InferenceMethod):
        im, rest = args[0], args[1:]
    else:
        im, rest = ExactInference(), args
    if len(rest) == 1 and isinstance(rest[0], InferenceState):
        inf = rest[0]
    elif len(rest) == 2:
        inf = InferenceState(rest[0], rest[1], evidence)
    else:
        raise TypeError("infer(im, bn, query; evidence) or infer(im, InferenceState)")
    return im.infer(inf)
