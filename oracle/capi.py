"""ctypes bindings of libbn_oracle.so (oracle/bn_oracle.c).  TEST INFRASTRUCTURE ONLY.

``net`` arguments are duck-typed: any object with numpy attributes
``card[int32 n] par_ptr[int32 n+1] par_idx[int32] cpt_off[int64 n+1] cpt[float64]``
(nodes in topological order, 0-based states, CPT node-state-fastest then first-parent-fastest).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbn_oracle.so")
_lib = None

REFERENCE_SCAN = 0  # Distributions.jl inverse-CDF scan in fp64
DEVICE_SPEC = 1     # same scan over 31-bit quantised thresholds (bit-exact with the CUDA kernels)

ANCESTRAL, REJECTION, WEIGHTED = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc; a few seconds)."""
    src = os.path.join(_HERE, "bn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libbn_oracle.so"] + (["-B"] if force else []))
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _net_args(net):
    card = np.ascontiguousarray(net.card, np.int32)
    par_ptr = np.ascontiguousarray(net.par_ptr, np.int32)
    par_idx = np.ascontiguousarray(net.par_idx, np.int32)
    if par_idx.size == 0:
        par_idx = np.zeros(1, np.int32)
    cpt_off = np.ascontiguousarray(net.cpt_off, np.int64)
    cpt = np.ascontiguousarray(net.cpt, np.float64)
    keep = (card, par_ptr, par_idx, cpt_off, cpt)
    args = (C.c_int32(card.size), _p(card, C.c_int32), _p(par_ptr, C.c_int32), _p(par_idx, C.c_int32),
            _p(cpt_off, C.c_int64), _p(cpt, C.c_double))
    return args, keep


def _evidence(net, evidence):
    if evidence is None:
        return None
    ev = np.ascontiguousarray(evidence, np.int8)
    assert ev.size == np.asarray(net.card).size
    return ev


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(t: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(t)))


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, np.uint32).copy()
    k = np.asarray(key, np.uint32).copy()
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def lw_infer(net, evidence, query, n, seed=0, first=0, mode=DEVICE_SPEC):
    """src/Inference/likelihood.jl:16-58 -> (unnormalised counts shaped by query cards (F order), sum w, sum w^2)."""
    args, keep = _net_args(net)
    ev = _evidence(net, evidence)
    q = np.ascontiguousarray(query, np.int32)
    shape = tuple(int(np.asarray(net.card)[i]) for i in q)
    counts = np.zeros(int(np.prod(shape)) if shape else 1, np.float64)
    stats = np.zeros(2, np.float64)
    rc = lib().orc_lw_infer(*args, _p(ev, C.c_int8), _p(q, C.c_int32), C.c_int32(q.size),
                            C.c_uint64(first), C.c_uint64(n), C.c_uint64(seed), C.c_int32(mode),
                            _p(counts, C.c_double), _p(stats, C.c_double))
    assert rc == 0
    return counts.reshape(shape, order="F"), float(stats[0]), float(stats[1])


def sample(net, n, evidence=None, kind=ANCESTRAL, seed=0, first=0, max_attempts=100, mode=DEVICE_SPEC):
    """src/sampling.jl:24-115 -> (states uint8 [n_nodes, n], weights [n], ok [n])."""
    args, keep = _net_args(net)
    ev = _evidence(net, evidence)
    nn = np.asarray(net.card).size
    states = np.zeros((nn, n), np.uint8)
    w = np.ones(n, np.float64)
    ok = np.ones(n, np.uint8)
    rc = lib().orc_sample(*args, _p(ev, C.c_int8), C.c_int32(kind), C.c_int32(max_attempts),
                          C.c_uint64(first), C.c_uint64(n), C.c_uint64(seed), C.c_int32(mode),
                          _p(states, C.c_uint8), _p(w, C.c_double), _p(ok, C.c_uint8))
    assert rc == 0
    return states, w, ok


def gibbs_infer(net, evidence, query, n_chains, nsamples, burn_in, thin, full=False, seed=0,
                first_chain=0, state=None, per_chain=False):
    """src/Inference/gibbs.jl:66-145 / :161-234 over n_chains chains -> (counts, final states[, per-chain counts])."""
    args, keep = _net_args(net)
    ev = _evidence(net, evidence)
    q = np.ascontiguousarray(query, np.int32)
    nn = np.asarray(net.card).size
    shape = tuple(int(np.asarray(net.card)[i]) for i in q)
    ncell = int(np.prod(shape)) if shape else 1
    counts = np.zeros(ncell, np.float64)
    init = 0
    if state is None:
        st = np.zeros((n_chains, nn), np.uint8)
    else:
        st = np.ascontiguousarray(state, np.uint8).copy()
        init = 1
    pcc = np.zeros((n_chains, ncell), np.float64) if per_chain else None
    rc = lib().orc_gibbs_infer(*args, _p(ev, C.c_int8), _p(q, C.c_int32), C.c_int32(q.size),
                               C.c_uint64(first_chain), C.c_uint64(n_chains), C.c_int64(nsamples),
                               C.c_int64(burn_in), C.c_int64(thin), C.c_int32(1 if full else 0),
                               C.c_uint64(seed), _p(st, C.c_uint8), C.c_int32(init),
                               _p(counts, C.c_double), _p(pcc, C.c_double))
    if rc != 0:
        raise ValueError("Burn in must be less than number of samples")  # src/Inference/gibbs.jl:70-71
    out = (counts.reshape(shape, order="F"), st)
    return out + (pcc,) if per_chain else out


def sweep_permutation(seed, sweep, m):
    perm = np.zeros(max(m, 1), np.int32)
    lib().orc_sweep_permutation(C.c_uint64(seed), C.c_uint64(sweep), C.c_int32(m), _p(perm, C.c_int32))
    return perm[:m]


def gibbs_sample(net, evidence, n_chains, nsamples, burn_in, thinning=0, variable_order=None, seed=0,
                 first_chain=0, init_state=None, log_domain=False):
    """src/gibbs.jl:289-374 -> states uint8 [n_nodes, n_chains*nsamples]."""
    args, keep = _net_args(net)
    ev = _evidence(net, evidence)
    nn = np.asarray(net.card).size
    vo = None if variable_order is None else np.ascontiguousarray(variable_order, np.int32)
    ist = None if init_state is None else np.ascontiguousarray(init_state, np.uint8)
    out = np.zeros((nn, n_chains * nsamples), np.uint8)
    rc = lib().orc_gibbs_sample(*args, _p(ev, C.c_int8), C.c_uint64(first_chain), C.c_uint64(n_chains),
                                C.c_int64(nsamples), C.c_int64(burn_in), C.c_int64(thinning),
                                _p(vo, C.c_int32), C.c_uint64(seed), _p(ist, C.c_uint8),
                                C.c_int32(1 if log_domain else 0), _p(out, C.c_uint8))
    if rc == -1:
        raise ValueError("bad gibbs_sample arguments")  # src/gibbs.jl:299-301
    if rc == -5:
        raise RuntimeError("Gibbs Sampler was unable to find an inital sample with non-zero probability")
    return out


# ---------------------------------------------------------------- factors (C, timed as CPU baseline)
def factor_product(dims1, a, dims2, b):
    """`*` = join(*, phi1, phi2): src/Factors/factors_dims.jl:185-234 -> (out_dims, out array F-order)."""
    a = np.asfortranarray(a, np.float64)
    b = np.asfortranarray(b, np.float64)
    d1 = np.ascontiguousarray(dims1, np.int32); c1 = np.ascontiguousarray(a.shape, np.int32)
    d2 = np.ascontiguousarray(dims2, np.int32); c2 = np.ascontiguousarray(b.shape, np.int32)
    if d1.size == 0: d1 = np.zeros(1, np.int32); c1 = np.ones(1, np.int32); nd1 = 0
    else: nd1 = d1.size
    if d2.size == 0: d2 = np.zeros(1, np.int32); c2 = np.ones(1, np.int32); nd2 = 0
    else: nd2 = d2.size
    ond = C.c_int32(0)
    od = np.zeros(nd1 + nd2 + 1, np.int32)
    oc = np.zeros(nd1 + nd2 + 1, np.int32)
    fa = a.ravel(order="F"); fb = b.ravel(order="F")
    rc = lib().orc_factor_product(C.c_int32(nd1), _p(d1, C.c_int32), _p(c1, C.c_int32), _p(fa, C.c_double),
                                  C.c_int32(nd2), _p(d2, C.c_int32), _p(c2, C.c_int32), _p(fb, C.c_double),
                                  C.byref(ond), _p(od, C.c_int32), _p(oc, C.c_int32), None)
    if rc == -2:
        raise ValueError("DimensionMismatch: Common dimensions must have same size")
    assert rc == 0
    nd = ond.value
    shape = tuple(int(x) for x in oc[:nd])
    out = np.zeros(int(np.prod(shape)) if nd else 1, np.float64)
    rc = lib().orc_factor_product(C.c_int32(nd1), _p(d1, C.c_int32), _p(c1, C.c_int32), _p(fa, C.c_double),
                                  C.c_int32(nd2), _p(d2, C.c_int32), _p(c2, C.c_int32), _p(fb, C.c_double),
                                  C.byref(ond), _p(od, C.c_int32), _p(oc, C.c_int32), _p(out, C.c_double))
    assert rc == 0
    return [int(x) for x in od[:nd]], out.reshape(shape, order="F")


def factor_sum(dims, a, sum_dims):
    """sum(phi, dims): src/Factors/factors_dims.jl:60-85,100."""
    a = np.asfortranarray(a, np.float64)
    d = np.ascontiguousarray(dims, np.int32); c = np.ascontiguousarray(a.shape, np.int32)
    s = np.ascontiguousarray(sum_dims, np.int32)
    ond = C.c_int32(0)
    od = np.zeros(d.size + 1, np.int32); oc = np.zeros(d.size + 1, np.int32)
    fa = a.ravel(order="F")
    sp = s if s.size else np.zeros(1, np.int32)
    rc = lib().orc_factor_sum(C.c_int32(d.size), _p(d, C.c_int32), _p(c, C.c_int32), _p(fa, C.c_double),
                              C.c_int32(s.size), _p(sp, C.c_int32), C.byref(ond), _p(od, C.c_int32),
                              _p(oc, C.c_int32), None)
    if rc == -3:
        raise ValueError("not a valid dimension")
    assert rc == 0
    nd = ond.value
    shape = tuple(int(x) for x in oc[:nd])
    out = np.zeros(int(np.prod(shape)) if nd else 1, np.float64)
    rc = lib().orc_factor_sum(C.c_int32(d.size), _p(d, C.c_int32), _p(c, C.c_int32), _p(fa, C.c_double),
                              C.c_int32(s.size), _p(sp, C.c_int32), C.byref(ond), _p(od, C.c_int32),
                              _p(oc, C.c_int32), _p(out, C.c_double))
    assert rc == 0
    return [int(x) for x in od[:nd]], out.reshape(shape, order="F")


def factor_normalize(a, p=1):
    """normalize!(phi; p): src/Factors/factors_dims.jl:45-57 (returns a normalised copy)."""
    a = np.array(a, np.float64, order="F", copy=True)
    flat = a.reshape(-1, order="F") if a.ndim else a.reshape(1)
    buf = np.ascontiguousarray(flat)
    rc = lib().orc_factor_normalize(C.c_int64(buf.size), _p(buf, C.c_double), C.c_int32(p))
    if rc != 0:
        raise ValueError(f"p = {p} is not supported")
    return buf.reshape(a.shape, order="F")


def statistics(net, data, n_rows=None, pitch=None):
    """src/DiscreteBayesNet/discrete_bayes_net.jl:151-172,221-231 -> int64 counts in the flat CPT layout
    (counts[cpt_off[i] + k + card_i * j]).  data: uint8 [n_nodes, pitch] node-major, 0-based states."""
    args, keep = _net_args(net)
    d = np.ascontiguousarray(data, np.uint8)
    n_nodes = int(np.asarray(net.card).size)
    if d.ndim == 2:
        assert d.shape[0] == n_nodes
        pitch = d.shape[1] if pitch is None else pitch
        n_rows = d.shape[1] if n_rows is None else n_rows
    out = np.zeros(int(np.asarray(net.cpt_off)[-1]), np.int64)
    rc = lib().orc_statistics(args[0], args[1], args[2], args[3], args[4], _p(d, C.c_uint8), C.c_uint64(n_rows),
                              C.c_uint64(pitch), _p(out, C.c_int64))
    assert rc == 0
    return out


def logpdf_table(net, data, n_rows=None, pitch=None) -> float:
    """logpdf(bn, df): src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-99 (row by row, node by
    node, same order of additions as the reference).  data: uint8 [n_nodes, pitch] node-major, 0-based states."""
    args, keep = _net_args(net)
    d = np.ascontiguousarray(data, np.uint8)
    if d.ndim == 2:
        assert d.shape[0] == int(np.asarray(net.card).size)
        pitch = d.shape[1] if pitch is None else pitch
        n_rows = d.shape[1] if n_rows is None else n_rows
    out = np.zeros(1, np.float64)
    rc = lib().orc_logpdf_table(*args, _p(d, C.c_uint8), C.c_uint64(n_rows), C.c_uint64(pitch), _p(out, C.c_double))
    assert rc == 0
    return float(out[0])


def loopy_infer(net, queries, evidences=None, nsamples=500, tol=1e-8, iters_for_convergence=6):
    """src/Inference/loopy_belief.jl:24-218 for a batch of (query node index, evidence row) instances (OpenMP over
    instances) -> (beliefs [n_inst, max_card] zero-padded, iterations [n_inst]).  The C restatement spells out the
    consequences of the reference's message aliasing; oracle.loopy_np keeps the aliasing itself; the tests require both
    to agree bit for bit."""
    args, keep = _net_args(net)
    q = np.ascontiguousarray(queries, np.int32)
    nn = np.asarray(net.card).size
    ev = None
    if evidences is not None:
        ev = np.ascontiguousarray(evidences, np.int8).reshape(q.size, nn)
    stride = int(np.asarray(net.card).max())
    out = np.zeros((q.size, stride), np.float64)
    its = np.zeros(q.size, np.int32)
    rc = lib().orc_loopy_infer(*args, _p(ev, C.c_int8), _p(q, C.c_int32), C.c_int64(q.size), C.c_int32(nsamples),
                               C.c_double(tol), C.c_int32(iters_for_convergence), C.c_int32(stride),
                               _p(out, C.c_double), _p(its, C.c_int32))
    if rc != 0:
        raise ValueError("iters_for_convergence must be >= 1 and nsamples >= 0")
    return out, its
