"""Table-producing samplers: rand(bn, n), rand(bn, n, evidence), weighted samples, gibbs_sample.

Mirrors src/sampling.jl:24-197 and src/gibbs.jl:289-462 for DiscreteBayesNets.  Every row is produced by
a CUDA thread (bn_sample / bn_gibbs_sample); this module only shapes the result into the reference's
DataFrame (one Int column per node, 1-based states).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import pandas as pd

from . import _lib
from ._lib import check, ptr
from .device_layout import device_net

_seed_counter = [0]


def _next_seed(seed):
    if seed is not None:
        return int(seed)
    _seed_counter[0] += 1
    return 0x5EED0000 + _seed_counter[0]


def _frame(lay, states: np.ndarray, extra: dict | None = None) -> pd.DataFrame:
    df = pd.DataFrame({n: states[i].astype(np.int64) + 1 for i, n in enumerate(lay.names)})
    for k, v in (extra or {}).items():
        df[k] = v
    return df


def _sample(bn, n, evidence, kind, max_attempts, seed, device, first_row=0):
    net = device_net(bn, device)
    lay = net.layout
    ev = lay.evidence_vector(evidence)
    n = int(n)
    states = np.zeros((lay.n, n), np.uint8)
    w = np.ones(n, np.float64)
    ok = np.ones(n, np.uint8)
    check(_lib.lib().bn_sample(_lib.h64(net.handle), ptr(ev, _lib.i8p), C.c_int32(kind), C.c_int32(max_attempts),
                               C.c_uint64(first_row), C.c_uint64(n), C.c_uint64(seed), ptr(states, _lib.u8p),
                               ptr(w, _lib.f64p), ptr(ok, _lib.u8p)))
    return lay, states, w, ok


class BayesNetSampler:
    """abstract type BayesNetSampler (src/sampling.jl:1-7)."""


@dataclass
class DirectSampler(BayesNetSampler):
    """Ancestral sampling (src/sampling.jl:50-63)."""


@dataclass
class RejectionSampler(BayesNetSampler):
    """src/sampling.jl:72-87: each row retried up to max_nsamples times."""
    evidence: dict = field(default_factory=dict)
    max_nsamples: int = 100


@dataclass
class LikelihoodWeightedSampler(BayesNetSampler):
    """src/sampling.jl:132-144: rows with evidence clamped plus a normalised weight column `potential`."""
    evidence: dict = field(default_factory=dict)


def get_weighted_dataframe(bn, nsamples: int, evidence: dict | None = None, seed=None, device=None) -> pd.DataFrame:
    """src/sampling.jl:153-174."""
    lay, states, w, _ = _sample(bn, nsamples, evidence or {}, _lib_kind("weighted"), 1, _next_seed(seed), device)
    with np.errstate(invalid="ignore", divide="ignore"):
        return _frame(lay, states, {"potential": w / w.sum()})


def get_weighted_sample(bn, evidence: dict | None = None, seed=None, device=None):
    """src/sampling.jl:102-121 -> (Assignment, weight)."""
    lay, states, w, _ = _sample(bn, 1, evidence or {}, _lib_kind("weighted"), 1, _next_seed(seed), device)
    return {n: int(states[i, 0]) + 1 for i, n in enumerate(lay.names)}, float(w[0])


def sample_weighted_dataframe(weighted_dataframe: pd.DataFrame, a: dict | None = None, u: float | None = None) -> dict:
    """sample_weighted_dataframe[!] (src/sampling.jl:126-127,182-197): one row drawn by a cumulative scan of the
    `potential` column against a uniform `u` (drawn here when not given); host-side, one row."""
    p = np.asarray(weighted_dataframe["potential"], np.float64)
    n = len(p)
    if n == 0:
        raise IndexError("BoundsError: empty weighted dataframe")  # p[1], :186
    u = float(np.random.default_rng().random()) if u is None else float(u)
    i, c = 0, p[0]
    while c < u and i < n - 1:
        i += 1
        c += p[i]
    a = {} if a is None else a
    for col in weighted_dataframe.columns:
        if col != "potential":
            v = weighted_dataframe[col].iloc[i]
            a[col] = int(v) if isinstance(v, (int, np.integer)) else v
    return a


def _lib_kind(name: str) -> int:
    return {"ancestral": 0, "rejection": 1, "weighted": 2}[name]


def rand(bn, *args, seed=None, device=None, **pairs):
    """rand(bn) -> Assignment | rand(bn, n) | rand(bn, n, evidence) | rand(bn, sampler[, n])
    (src/sampling.jl:17-41,61-63,89-93).  Keyword pairs play the role of `:c=>1` pairs."""
    sampler, n, evidence = DirectSampler(), None, None
    for a in args:
        if isinstance(a, BayesNetSampler):
            sampler = a
        elif isinstance(a, dict):
            evidence = a
        else:
            n = int(a)
    if pairs:
        evidence = dict(evidence or {}, **pairs)
    if evidence is not None and isinstance(sampler, DirectSampler):
        sampler = RejectionSampler(evidence, 100)  # src/sampling.jl:89-90
    seed = _next_seed(seed)
    single = n is None
    n = 1 if single else n

    if isinstance(sampler, GibbsSampler):
        df = gibbs_sample(bn, n, sampler.burn_in, thinning=sampler.thinning, consistent_with=sampler.evidence,
                          variable_order=sampler.variable_order, time_limit=sampler.time_limit,
                          error_if_time_out=sampler.error_if_time_out, initial_sample=sampler.initial_sample,
                          max_cache_size=sampler.max_cache_size, seed=seed, device=device)
    elif isinstance(sampler, LikelihoodWeightedSampler):
        df = get_weighted_dataframe(bn, n, sampler.evidence, seed=seed, device=device)
    elif isinstance(sampler, RejectionSampler):
        lay, states, _, ok = _sample(bn, n, sampler.evidence, _lib_kind("rejection"), sampler.max_nsamples, seed, device)
        if not ok.all():
            # the reference returns an emptied Assignment and then fails on a[name] (src/sampling.jl:87,33-36)
            raise KeyError(f"rejection sampling found no sample consistent with the evidence in "
                           f"{sampler.max_nsamples} tries for {int((ok == 0).sum())} of {n} rows")
        df = _frame(lay, states)
    else:
        lay, states, _, _ = _sample(bn, n, None, _lib_kind("ancestral"), 1, seed, device)
        df = _frame(lay, states)
    if single:
        return {c: (int(df[c].iloc[0]) if c != "potential" else float(df[c].iloc[0])) for c in df.columns}
    return df


# ---------------------------------------------------------------------------------------------
def gibbs_sample(bn, nsamples: int, burn_in: int, *, thinning: int = 0, consistent_with: dict | None = None,
                 variable_order: Optional[Sequence[str]] = None, time_limit: Optional[int] = None,
                 error_if_time_out: bool = True, initial_sample: Optional[dict] = None,
                 max_cache_size: Optional[int] = None, n_chains: int = 1, seed=None, device=None) -> pd.DataFrame:
    """gibbs_sample (src/gibbs.jl:289-374).  Returns nsamples rows per chain (n_chains=1 is the reference's
    single chain); evidence columns are appended as Float64 like the reference (:368-373).

    time_limit is validated but never triggers: the device loop has no wall clock, and the reference only
    uses it to cut a slow CPU loop short (:218-220).  max_cache_size is accepted and unused: the string-keyed
    conditional cache (:43-71) is a CPU optimisation."""
    # argument validation, same order and messages as src/gibbs.jl:299-324
    if not nsamples > 0:
        raise ValueError("nsamples parameter less than 1")
    if not burn_in >= 0:
        raise ValueError("Negative burn_in parameter")
    if not thinning >= 0:
        raise ValueError("Negative thinning parameter")
    names = bn.names()
    if variable_order is not None:
        for name in names:
            if name not in variable_order:
                raise ValueError("Gibbs sample variable_order must contain all variables in the Bayes Net")
        for name in variable_order:
            if name not in names:
                raise ValueError("Gibbs sample variable_order contains a variable not in the Bayes Net")
    if time_limit is not None and not time_limit > 0:
        raise ValueError(f"Invalid time_limit specified ({time_limit})")
    consistent_with = dict(consistent_with or {})
    if initial_sample is not None:
        for name in names:
            if name not in initial_sample:
                raise ValueError("Gibbs sample initial_sample must be an assignment with all variables in the Bayes Net")
        for name, v in consistent_with.items():
            if initial_sample[name] != v:
                raise ValueError("Gibbs sample initial_sample was inconsistent with consistent_with")
        if not bn.pdf(initial_sample) > 0:
            raise ValueError("Gibbs sample initial_sample has a pdf value of zero")

    net = device_net(bn, device)
    lay = net.layout
    ev = lay.evidence_vector(consistent_with)
    free = [n for n in lay.names if ev[lay.index[n]] < 0]
    vo = None
    if variable_order is not None:
        pos = {n: i for i, n in enumerate(free)}
        vo = np.array([pos[n] for n in variable_order if n in pos], np.int32)
    init = None
    if initial_sample is not None:
        row = np.array([int(initial_sample[n]) - 1 for n in lay.names], np.uint8)
        init = np.ascontiguousarray(np.tile(row, (int(n_chains), 1)))
    out = np.zeros((lay.n, int(n_chains) * int(nsamples)), np.uint8)
    check(_lib.lib().bn_gibbs_sample(_lib.h64(net.handle), ptr(ev, _lib.i8p), C.c_uint64(0), C.c_uint64(n_chains),
                                     C.c_int64(nsamples), C.c_int64(burn_in), C.c_int64(thinning),
                                     ptr(vo, _lib.i32p), C.c_uint64(_next_seed(seed)), ptr(init, _lib.u8p),
                                     ptr(out, _lib.u8p)))
    cols = {n: out[lay.index[n]].astype(np.int64) + 1 for n in free}
    for n in consistent_with:  # evidence columns as Float64 (:368-373)
        if n in lay.index:
            cols[n] = np.full(out.shape[1], float(consistent_with[n]))
    return pd.DataFrame(cols)


@dataclass
class GibbsSampler(BayesNetSampler):
    """src/gibbs.jl:410-433."""
    evidence: dict = field(default_factory=dict)
    burn_in: int = 100
    thinning: int = 0
    variable_order: Optional[Sequence[str]] = None
    time_limit: Optional[int] = None
    error_if_time_out: bool = True
    initial_sample: Optional[dict] = None
    max_cache_size: Optional[int] = None
