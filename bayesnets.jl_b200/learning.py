"""Sufficient statistics, Dirichlet-prior Bayesian score and MLE parameter fit for discrete networks -- the step
after sampling in a learn-from-samples loop.

Mirrors, name for name:
  statistics(...)                 src/DiscreteBayesNet/discrete_bayes_net.jl:151-172 (one node), :221-231 (all nodes),
                                  :233-257 (DataFrame front ends: parents = inneighbors(dag, i), i.e. ASCENDING node
                                  index, ncategories inferred from the data = max value per column)
  UniformPrior / BDeuPrior        src/DiscreteBayesNet/dirichlet_priors.jl
  bayesian_score_component(s)     src/DiscreteBayesNet/structure_scoring.jl:17-42,75-94,133-160
  bayesian_score                  src/DiscreteBayesNet/structure_scoring.jl:95-125
  logpdf(bn, df) / pdf(bn, df)    src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-104 (the data
                                  log-likelihood: the table's counts dotted with the log CPTs, bn_logpdf[_dev])
  fit(DiscreteBayesNet, data, edges)   src/learning.jl:1-64 + src/CPDs/categorical_cpd.jl:109-152 (MLE per parental
                                  instantiation; an instantiation never seen gets the uniform Categorical)

The counting itself -- one pass over the n_nodes x n_rows table for ALL nodes -- is the CUDA kernel behind
bn_statistics[_dev] (csrc/stats.cu); `statistics_device` takes the uint8 node-major table that bn_sample_dev
leaves in HBM, so rand(bn, n) -> statistics -> fit never touches the host.  The scores are a few thousand
log-gammas of the resulting counts and stay on the host (math.lgamma), as cheap there as the reference's.
States are 1-based in DataFrames / matrices (as in Julia) and 0-based uint8 in device tables.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import check, ptr
from .bayes_net import BayesNet, CategoricalCPD
from .device_layout import DeviceLayout, DeviceNet, default_device


# ------------------------------------------------------------------------------------------------ structure
def _structure_layout(parent_list: Sequence[Sequence[int]], ncategories: Sequence[int], names=None) -> DeviceLayout:
    """A structure-only DeviceLayout (0-based parents in the given order; CPT values are never read by the kernel)."""
    n = len(parent_list)
    card = np.asarray(ncategories, np.int32)
    par_ptr, par_idx, cpt_off = [0], [], [0]
    for i in range(n):
        q = 1
        for p in parent_list[i]:
            if not 0 <= p < n or p == i:
                raise ValueError(f"node {i + 1}: parent index {p + 1} out of range")
            par_idx.append(int(p))
            q *= int(card[p])
        par_ptr.append(len(par_idx))
        cpt_off.append(cpt_off[-1] + int(card[i]) * q)
    names = list(names) if names is not None else [f"X{i + 1}" for i in range(n)]
    return DeviceLayout(names, {nm: i for i, nm in enumerate(names)}, card, np.asarray(par_ptr, np.int32),
                        np.asarray(par_idx, np.int32), np.asarray(cpt_off, np.int64), np.zeros(cpt_off[-1]))


def _split(lay: DeviceLayout, counts: np.ndarray) -> list:
    out = []
    for i in range(lay.n):
        q = 1
        for p in lay.parents(i):
            q *= int(lay.card[p])
        out.append(np.asarray(counts[lay.cpt_off[i]:lay.cpt_off[i + 1]], np.int64).reshape((int(lay.card[i]), q), order="F"))
    return out


def infer_number_of_instantiations(col) -> int:
    """src/CPDs/utils.jl: the number of instantiations of a discrete column = its maximum value."""
    col = np.asarray(col)
    return int(col.max()) if col.size else 0


def _matrix_from_frame(data, names) -> np.ndarray:
    """DataFrame (columns = node names, 1-based states) -> int matrix n x m like `Matrix{Int}(data)'`."""
    return np.stack([np.asarray(data[nm], np.int64) for nm in names]) if len(names) else np.zeros((0, 0), np.int64)


def _to_u8(datamat: np.ndarray, card: np.ndarray) -> np.ndarray:
    d = np.asarray(datamat)
    if d.ndim != 2:
        raise ValueError("data must be an n_nodes x n_samples matrix")
    if d.size and (d.min() < 1 or (d > np.asarray(card)[:, None]).any()):
        raise IndexError("BoundsError: data holds a state outside 1:ncategories")  # sparse(...) would throw
    if (np.asarray(card) > 255).any():
        raise NotImplementedError("more than 255 states per variable")
    return np.ascontiguousarray(d - 1, np.uint8)


# ------------------------------------------------------------------------------------------------ statistics
def _counts_host_table(lay: DeviceLayout, table_u8: np.ndarray, device) -> np.ndarray:
    net = DeviceNet(lay, default_device() if device is None else device)
    try:
        out = np.zeros(int(lay.cpt_off[-1]), np.int64)
        check(_lib.lib().bn_statistics(_lib.h64(net.handle), ptr(table_u8, _lib.u8p), C.c_uint64(table_u8.shape[1]),
                                       ptr(out, _lib.i64p)))
        return out
    finally:
        net.close()


def statistics(*args, device=None):
    """statistics(targetind, parents, ncategories, data) -> r x q matrix            (:151-172, 1-based indices)
    statistics(parent_list, ncategories, data)         -> list of r_i x q_i matrices (:221-231)
    statistics(bn, data) / statistics(bn, target, data) with a DataFrame             (:233-257)"""
    if isinstance(args[0], BayesNet):
        bn = args[0]
        names = bn.names()
        data = args[-1]
        if len(data.columns) != len(names):
            raise DimensionMismatch_(len(names), len(data.columns))
        # Matrix{Int}(data)': columns are taken BY POSITION (column i belongs to bn.cpds[i]), as the reference does
        mat = np.stack([np.asarray(data.iloc[:, i], np.int64) for i in range(len(names))])
        ncat = [infer_number_of_instantiations(mat[i]) for i in range(len(names))]
        idx = {nm: i for i, nm in enumerate(names)}
        plist = [sorted(idx[p] for p in bn.parents(nm)) for nm in names]  # inneighbors(dag, i): ascending
        if len(args) == 3:
            t = idx[args[1]]
            return statistics(t + 1, [p + 1 for p in plist[t]], ncat, mat, device=device)
        return statistics([[p + 1 for p in ps] for ps in plist], ncat, mat, device=device)
    if isinstance(args[0], (int, np.integer)):
        target, parents, ncat, mat = args
        plist = [[] for _ in range(len(ncat))]
        plist[int(target) - 1] = [int(p) - 1 for p in parents]
        return _statistics_lists(plist, ncat, mat, device)[int(target) - 1]
    parent_list, ncat, mat = args
    return _statistics_lists([[int(p) - 1 for p in ps] for ps in parent_list], ncat, mat, device)


def _topological_order(plist):
    """Kahn's algorithm over 0-based parent lists -> order (list of old indices) or None when cyclic."""
    n = len(plist)
    indeg = [len(ps) for ps in plist]
    kids = [[] for _ in range(n)]
    for i, ps in enumerate(plist):
        for p in ps:
            kids[p].append(i)
    ready = [i for i in range(n) if indeg[i] == 0]
    order = []
    while ready:
        u = ready.pop(0)
        order.append(u)
        for c in kids[u]:
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(c)
    return order if len(order) == n else None


def _statistics_lists(plist, ncat, mat, device):
    """All nodes in one device pass.  bn_net_create wants parents before children, `statistics` does not care
    ("No acyclicity checking is done", :189), so the rows are relabelled into a topological order first; a cyclic
    parent list falls back to one pass per node (target last)."""
    n = len(plist)
    mat = np.asarray(mat)
    order = _topological_order(plist)
    if order is None:
        out = []
        for t in range(n):
            single = [[] for _ in range(n)]
            single[t] = plist[t]
            out.append(_statistics_lists(single, ncat, mat, device)[t])
        return out
    pos = {old: new for new, old in enumerate(order)}
    lay = _structure_layout([[pos[p] for p in plist[old]] for old in order], [ncat[old] for old in order])
    Ns = _split(lay, _counts_host_table(lay, _to_u8(mat[order], lay.card), device))
    return [Ns[pos[i]] for i in range(n)]


def DimensionMismatch_(n, m):
    return _lib.DimensionMismatch(f"statistics' dag and data must be of the same dimension, {n} ≠ {m}")


def statistics_device(bn_or_net, states_dev_ptr: int, n_rows: int, pitch: int | None = None, out_dev_ptr: int | None = None):
    """Counts of a DEVICE-resident uint8 node-major table (what bn_sample_dev / rand(..., device table) writes) under
    the structure of `bn_or_net` (parents in cpd.parents order, cards of the network): returns (layout, counts) with
    counts in the flat CPT layout, or enqueues into `out_dev_ptr` (int64[cpt_len], zeroed by the call) and returns
    (layout, None) without synchronising."""
    from .device_layout import device_net
    net = bn_or_net if isinstance(bn_or_net, DeviceNet) else device_net(bn_or_net)
    lay = net.layout
    pitch = n_rows if pitch is None else pitch
    if out_dev_ptr is not None:
        check(_lib.lib().bn_statistics_dev(_lib.h64(net.handle), C.c_void_p(states_dev_ptr), C.c_uint64(n_rows),
                                           C.c_uint64(pitch), C.c_void_p(out_dev_ptr)))
        return lay, None
    n_bins = int(lay.cpt_off[-1])
    out = _lib.DeviceBuffer(n_bins * 8, net.device)  # bn_dev_alloc: no array library on the product path
    check(_lib.lib().bn_statistics_dev(_lib.h64(net.handle), C.c_void_p(states_dev_ptr), C.c_uint64(n_rows),
                                       C.c_uint64(pitch), C.c_void_p(out.ptr)))
    counts = out.download(np.int64, n_bins)
    out.close()
    return lay, counts


# ------------------------------------------------------------------------------------------------ priors + score
class DirichletPrior:
    def get(self, r: int, q: int) -> np.ndarray:
        raise NotImplementedError


@dataclass(frozen=True)
class UniformPrior(DirichletPrior):
    """All pseudo-counts equal α (default: the K2 prior α = 1)."""
    alpha: float = 1.0

    def get(self, r, q):
        return np.full((r, q), float(self.alpha))


@dataclass(frozen=True)
class BDeuPrior(DirichletPrior):
    """α_ijk = x / (q_i r_i): equal scores for Markov-equivalent structures."""
    x: float = 1.0

    def get(self, r, q):
        return np.full((r, q), float(self.x) / (r * q))


_lgamma = np.vectorize(math.lgamma, otypes=[np.float64])


def _score_from_counts(N: np.ndarray, alpha: np.ndarray) -> float:
    """structure_scoring.jl:40-41 on the r x q count matrix."""
    N = np.asarray(N, np.float64)
    a0 = alpha.sum(axis=0)
    return float(_lgamma(alpha + N).sum() - _lgamma(alpha).sum() + _lgamma(a0).sum() - _lgamma(a0 + N.sum(axis=0)).sum())


def bayesian_score_component(i, parents, ncategories, data, prior: DirichletPrior = UniformPrior(), device=None) -> float:
    """Score of variable i (1-based) with the given parent indices (1-based) on the n x m data matrix."""
    N = statistics(int(i), list(parents), list(ncategories), data, device=device)
    return _score_from_counts(N, prior.get(*N.shape))


def bayesian_score_components(*args, device=None) -> list:
    """bayesian_score_components(parent_list, ncategories, data, prior) | (bn, data, prior): one counting pass."""
    if isinstance(args[0], BayesNet):
        bn, data = args[0], args[1]
        prior = args[2] if len(args) > 2 else UniformPrior()
        Ns = statistics(bn, data, device=device)
    else:
        parent_list, ncat, mat = args[:3]
        prior = args[3] if len(args) > 3 else UniformPrior()
        Ns = statistics(parent_list, ncat, mat, device=device)
    return [_score_from_counts(N, prior.get(*N.shape)) for N in Ns]


def bayesian_score(*args, device=None) -> float:
    """bayesian_score(parent_list, ncategories, data, prior) | bayesian_score(bn, data[, prior])."""
    tot = 0.0
    for s in bayesian_score_components(*args, device=device):
        tot += s
    return tot


# ------------------------------------------------------------------------------------------------ fit
def _cpds_from_counts(names, plist, card, Ns) -> list:
    cpds = []
    for i, N in enumerate(Ns):
        r = int(card[i])
        dists = []
        for j in range(N.shape[1]):
            tot = int(N[:, j].sum())
            # fit_mle(Categorical, r, arr) = counts / total; unseen instantiation -> Categorical(r) (uniform)
            dists.append(N[:, j].astype(np.float64) / tot if tot else np.full(r, 1.0 / r))
        cpds.append(CategoricalCPD(names[i], [names[p] for p in plist[i]], [int(card[p]) for p in plist[i]], dists))
    return cpds


def fit(data, edges=(), names=None, ncategories=None, device=None) -> BayesNet:
    """fit(DiscreteBayesNet, data, edges): MLE CPTs of the DAG given as (parent, child) name pairs (src/learning.jl:1-64
    with DiscreteCPD, src/CPDs/categorical_cpd.jl:109-152).  `data`: DataFrame with 1-based integer states."""
    names = list(data.columns) if names is None else list(names)
    idx = {nm: i for i, nm in enumerate(names)}
    if isinstance(edges, tuple) and len(edges) == 2 and not isinstance(edges[0], (tuple, list)):
        edges = (edges,)  # a single :A => :B pair
    plist = [[] for _ in names]
    for pa, ch in edges:
        plist[idx[ch]].append(idx[pa])
    plist = [sorted(ps) for ps in plist]  # parents = names[inneighbors(dag, i)] (src/learning.jl:8)
    if _topological_order(plist) is None:
        raise ValueError("BayesNet graph is non-acyclic!")
    mat = _matrix_from_frame(data, names)
    ncat = [infer_number_of_instantiations(mat[i]) for i in range(len(names))] if ncategories is None else list(ncategories)
    Ns = _statistics_lists(plist, ncat, mat, device)
    return BayesNet(_cpds_from_counts(names, plist, ncat, Ns))


def fit_from_device_table(bn_structure: BayesNet, states_dev_ptr: int, n_rows: int, pitch: int | None = None) -> BayesNet:
    """MLE re-fit of `bn_structure`'s CPTs from a device-resident sample table (same structure and cards)."""
    lay, counts = statistics_device(bn_structure, states_dev_ptr, n_rows, pitch)
    plist = [lay.parents(i) for i in range(lay.n)]
    return BayesNet(_cpds_from_counts(lay.names, plist, lay.card, _split(lay, counts)))


# ------------------------------------------------------------------------------------------------ data likelihood
def logpdf(bn: BayesNet, data, device=None) -> float:
    """logpdf(bn, assignment) (src/bayes_nets.jl:322-328) or logpdf(bn, df::DataFrame)
    (ProbabilisticGraphicalModels.jl:83-99): sum over rows of sum over nodes of log P(x_i | pa_i).  Columns are taken
    BY NAME (`a[name] = df[i, name]`, :91); extra columns are ignored, a missing one raises KeyError.  A value outside
    1:ncategories gives pdf 0 for its own node (-> -Inf) but is a BoundsError as a parent (categorical_cpd.jl:36-46)."""
    if isinstance(data, dict):
        return bn.logpdf(data)
    from .device_layout import device_net
    net = device_net(bn, default_device() if device is None else device)
    lay = net.layout
    mat = _matrix_from_frame(data, lay.names)
    if mat.shape[0] == 0 or mat.shape[1] == 0:
        return 0.0
    card = np.asarray(lay.card)
    bad = (mat < 1) | (mat > card[:, None])
    if bad.any():
        is_parent = np.zeros(lay.n, bool)
        is_parent[np.asarray(lay.par_idx[:lay.par_ptr[-1]], np.int64)] = True
        if bad[is_parent].any():
            raise IndexError("BoundsError: a parent's value lies outside 1:ncategories")
        return -math.inf
    if (card > 255).any():
        raise NotImplementedError("more than 255 states per variable")
    table = np.ascontiguousarray(mat - 1, np.uint8)
    out = np.zeros(1, np.float64)
    check(_lib.lib().bn_logpdf(_lib.h64(net.handle), ptr(table, _lib.u8p), C.c_uint64(table.shape[1]), ptr(out, _lib.f64p)))
    return float(out[0])


def pdf(bn: BayesNet, data, device=None) -> float:
    """pdf(bn, assignment) (src/bayes_nets.jl:311-317) or pdf(bn, df) = exp(logpdf(bn, df)) (:104)."""
    if isinstance(data, dict):
        return bn.pdf(data)
    return math.exp(logpdf(bn, data, device=device))


def logpdf_device(bn_or_net, states_dev_ptr: int, n_rows: int, pitch: int | None = None, out_dev_ptr: int | None = None):
    """Log-likelihood of a DEVICE-resident uint8 node-major table (what bn_sample_dev writes) under the network:
    returns the value, or enqueues it into `out_dev_ptr` (one double) and returns None without synchronising."""
    from .device_layout import device_net
    net = bn_or_net if isinstance(bn_or_net, DeviceNet) else device_net(bn_or_net)
    pitch = n_rows if pitch is None else pitch
    if out_dev_ptr is not None:
        check(_lib.lib().bn_logpdf_dev(_lib.h64(net.handle), C.c_void_p(states_dev_ptr), C.c_uint64(n_rows),
                                       C.c_uint64(pitch), C.c_void_p(out_dev_ptr)))
        return None
    out = _lib.DeviceBuffer(8, net.device)  # bn_dev_alloc: no array library on the product path
    check(_lib.lib().bn_logpdf_dev(_lib.h64(net.handle), C.c_void_p(states_dev_ptr), C.c_uint64(n_rows),
                                   C.c_uint64(pitch), C.c_void_p(out.ptr)))
    v = float(out.download(np.float64, 1)[0])
    out.close()
    return v


# ------------------------------------------------------------------------------------------------ count / table
def count(bn: BayesNet, *args, device=None):
    """Base.count(bn, name, data) -> DataFrame of the observed (parents..., name) assignments and their counts, rows in
    order of first appearance (`unique(t)`), columns `parents(bn, name)..., name, count`; count(bn, data) -> one such
    frame per node (src/DiscreteBayesNet/discrete_bayes_net.jl:101-121).  The tally is statistics(bn, name, data) on the
    device (the reference compares every unique row with every row); the host only lists the rows that occur."""
    import pandas as pd
    if len(args) == 1:
        return [count(bn, nm, args[0], device=device) for nm in bn.names()]
    name, data = args
    varnames = list(bn.parents(name)) + [name]
    t = data[varnames]
    tu = t.drop_duplicates(ignore_index=True)
    if len(tu) == 0:
        return tu.assign(count=np.zeros(0, np.int64))
    N = statistics(bn, name, data, device=device)  # r x q, parental instantiations over the parents in ASCENDING node order
    names = bn.names()
    order = sorted(bn.parents(name), key=names.index)
    ncat = {p: infer_number_of_instantiations(data.iloc[:, names.index(p)]) for p in order}
    j = np.zeros(len(tu), np.int64)
    stride = 1
    for p in order:
        j += stride * (np.asarray(tu[p], np.int64) - 1)
        stride *= ncat[p]
    return tu.assign(count=N[np.asarray(tu[name], np.int64) - 1, j])


def table(bn: BayesNet, name, assignment: dict | None = None, /, **kw):
    """table(bn, name[, assignment]) -> DataFrame of the node's CPT: columns `parents..., name, potential`, first parent
    fastest, the node's own state slowest (src/DiscreteBayesNet/discrete_bayes_net.jl:49-91); with an assignment, only
    the rows consistent with it (:88-91).  Host-side formatting of the CPT the device tables are built from."""
    import pandas as pd
    cpd = bn.get(name)
    pn = [int(x) for x in cpd.parental_ncategories]
    r = cpd.ncategories()
    q = int(np.prod(pn)) if pn else 1
    cols = {}
    inner = 1
    for p, c in zip(cpd.parents, pn):
        cols[p] = np.tile(np.repeat(np.arange(1, c + 1), inner), (q // (inner * c)) * r)
        inner *= c
    cols[name] = np.repeat(np.arange(1, r + 1), q)
    cols["potential"] = np.array([cpd.distributions[jj][k] for k in range(r) for jj in range(q)], np.float64)
    df = pd.DataFrame(cols)
    a = dict(assignment or {}, **kw)
    for k, v in a.items():
        if k in df.columns:
            df = df[df[k] == v]
    return df.reset_index(drop=True)
