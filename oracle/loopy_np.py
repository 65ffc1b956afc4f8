"""numpy restatement of infer(::LoopyBelief) (src/Inference/loopy_belief.jl:24-218).  TEST INFRASTRUCTURE ONLY.

The reference passes messages around as Julia arrays *by reference*, and three of those references matter
for the numbers it produces (none is documented, all are reproduced here because numpy arrays alias the
same way Julia arrays do):

  * ``pi = phi.potential`` for a parentless node (:125) is the node's own factor, which lives across
    iterations; ``px = pi; px .*= ...; normalize!(px, 1)`` (:170-175) therefore rewrites the prior of a
    root node in place on every iteration, once per child.
  * ``new_pis[(nn, ch)] = px`` (:177) stores the *same* array for every child of ``nn``, so every child
    receives the message left after the last child's update, not its own.
  * after the second iteration both message dictionaries hold the root's array, so a root's children read
    the message computed in the *current* iteration and its change ``norm(px - pis[..], Inf)`` is 0.

Arithmetic order follows the calls the reference makes:
  broadcast(*, phi, dims, values)  -> left fold ``((phi * v1) * v2) ...`` per element (factors_dims.jl:131-165)
  sum!(phi, dims)                  -> sequential accumulation in column-major order (factors_dims.jl:87-98)
  normalize!(::Factor)             -> ``./= sum(abs, .)`` (factors_dims.jl:45-57)
  normalize!(::Vector, 1)          -> ``rmul!(v, inv(norm(v, 1)))`` (LinearAlgebra; two-step scaling below 1/floatmax)
  norm(v) / norm(v, Inf)           -> LinearAlgebra generic_norm2 / generic_normInf for short vectors

States are 0-based here.  Pinned by the four LoopyBelief posteriors of test/test_inference.jl:45-63,94-110
(tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

from .factors_np import FlatNet

_FLOATMAX = np.finfo(np.float64).max
_DELTA = 1.0 / np.nextafter(np.inf, 0.0)          # inv(prevfloat(typemax(Float64)))
_EPS_DELTA = np.finfo(np.float64).eps / _DELTA


def _seq_sum(v) -> float:
    """Left-to-right sum (Base mapreduce for < 16 elements; the convention of factors_np.NpFactor.sum)."""
    acc = float(v[0])
    for t in v[1:]:
        acc = acc + float(t)
    return acc


def _norm1(v) -> float:
    return _seq_sum(np.abs(v))


def _norm_inf(v) -> float:
    """generic_normInf: max with NaN propagation."""
    m = 0.0
    for t in v:
        a = abs(float(t))
        if a != a:
            return math.nan
        m = max(m, a)
    return m


def _norm2(v) -> float:
    """generic_norm2 (LinearAlgebra/src/generic.jl): plain sum of squares unless it would under/overflow."""
    maxabs = _norm_inf(v)
    if maxabs == 0.0 or math.isinf(maxabs):
        return maxabs
    if maxabs != maxabs:
        return math.nan
    with np.errstate(over="ignore", under="ignore"):
        mm = np.float64(maxabs) * np.float64(maxabs)
        if np.isfinite(np.float64(len(v)) * mm) and mm != 0.0:
            acc = 0.0
            for t in v:
                acc = acc + float(t) * float(t)
            return math.sqrt(acc)
        acc = 0.0
        for t in v:
            r = abs(float(t)) / maxabs
            acc = acc + r * r
        return maxabs * math.sqrt(acc)


def _normalize1_inplace(v: np.ndarray) -> None:
    """LinearAlgebra.normalize!(v, 1) -> __normalize!(v, norm(v, 1))."""
    nrm = _norm1(v)
    with np.errstate(all="ignore"):
        if nrm >= _DELTA:
            v *= np.float64(1.0) / np.float64(nrm)
        else:
            v *= _EPS_DELTA
            v *= np.float64(1.0) / (np.float64(nrm) * _EPS_DELTA)


def _factor_normalize_inplace(v: np.ndarray) -> None:
    """normalize!(::Factor): potential ./= sum(abs, potential)."""
    total = _norm1(v.reshape(-1, order="F"))
    with np.errstate(all="ignore"):
        v /= np.float64(total)


def _broadcast_mul(phi: np.ndarray, axes, values) -> np.ndarray:
    """broadcast(*, phi, dims, values): a copy of phi with each vector multiplied in along its axis,
    one after the other in the order given (n-ary `*` is a left fold)."""
    out = phi.copy(order="F")
    for ax, v in zip(axes, values):
        shape = [1] * out.ndim
        shape[ax] = len(v)
        out = out * np.asarray(v).reshape(shape)
    return out


def _sum_axes(a: np.ndarray, axes) -> np.ndarray:
    """sum!(phi, dims) + dropdims: sequential accumulation over the reduced axes in column-major order."""
    axes = sorted(axes)
    if not axes:
        return a
    moved = np.moveaxis(a, axes, range(len(axes)))
    flat = moved.reshape((-1,) + moved.shape[len(axes):], order="F")
    acc = flat[0].copy()
    for t in range(1, flat.shape[0]):
        acc = acc + flat[t]
    return acc


def _fold_mul(vs):
    vs = list(vs)
    out = vs[0]
    for v in vs[1:]:
        out = out * v
    return out


def loopy_infer(net: FlatNet, query: int, evidence=None, nsamples: int = 500, tol: float = 1e-8,
                iters_for_convergence: int = 6, return_iters: bool = False):
    """P(query | evidence) by loopy belief propagation, exactly as the reference computes it.

    net       FlatNet-like (card, par_ptr, par_idx, cpt_off, cpt), nodes in topological order
    query     node index
    evidence  int8[n_nodes], -1 = free, else the 0-based observed state (or None)
    """
    n = len(net.card)
    card = [int(c) for c in net.card]
    ev = [-1] * n if evidence is None else [int(e) for e in evidence]
    parents = [[int(p) for p in net.par_idx[net.par_ptr[i]:net.par_ptr[i + 1]]] for i in range(n)]
    children = [[j for j in range(n) if i in parents[j]] for i in range(n)]          # ascending, like outneighbors
    factors = [np.array(net.cpt[net.cpt_off[i]:net.cpt_off[i + 1]], np.float64).reshape(
        [card[i]] + [card[p] for p in parents[i]], order="F") for i in range(n)]      # Factor(bn, nn) :40
    ev_lambda = {}
    for i in range(n):
        if ev[i] >= 0:
            z = np.zeros(card[i]); z[ev[i]] = 1.0                                     # :17-22
            ev_lambda[i] = z

    lambdas, pis = {}, {}
    for i in range(n):                                                                 # :69-85
        for ch in children[i]:
            lambdas[(ch, i)] = np.full(card[i], 1.0 / card[i])
            pis[(i, ch)] = ev_lambda[i] if i in ev_lambda else np.full(card[i], 1.0 / card[i])
    new_lambdas = {}
    new_pis = {k: v.copy() for k, v in pis.items()}                                    # deepcopy :90

    change_index = 0
    change_per_iter = [math.inf] * iters_for_convergence
    iters_done = 0
    for _ in range(nsamples):                                                          # :98
        iters_done += 1
        max_change = -math.inf
        for i in range(n):
            phi = factors[i]
            if i in ev_lambda:                                                         # :114-121
                lam = ev_lambda[i]
            elif not children[i]:
                lam = np.ones(card[i])
            else:
                lam = _fold_mul(lambdas[(ch, i)] for ch in children[i])
            if not parents[i]:                                                         # :125-133
                pi = phi                                                               # alias of the node's factor
            else:
                ft = _broadcast_mul(phi, range(1, 1 + len(parents[i])), [pis[(p, i)] for p in parents[i]])
                pi = _sum_axes(ft, range(1, 1 + len(parents[i])))
            for j, pa in enumerate(parents[i]):                                        # :136-160
                other = [jj for jj in range(len(parents[i])) if jj != j]
                lx = _broadcast_mul(phi, [1 + jj for jj in other], [pis[(parents[i][jj], i)] for jj in other])
                lx = _sum_axes(lx, [1 + jj for jj in other])                           # [nn, pa]
                lx = lx * lam.reshape(-1, 1)
                lx = _sum_axes(lx, [0])
                _factor_normalize_inplace(lx)
                new_lambdas[(i, pa)] = lx
                delta = _norm2(lx - lambdas[(i, pa)])
                if delta > max_change:
                    max_change = delta
            if i not in ev_lambda:                                                     # :163-185
                for ch in children[i]:
                    other_ch = [c for c in children[i] if c != ch]
                    px = pi
                    if other_ch:
                        px *= _fold_mul(lambdas[(c, i)] for c in other_ch)             # in place
                    _normalize1_inplace(px)
                    new_pis[(i, ch)] = px
                    with np.errstate(invalid="ignore"):
                        delta = _norm_inf(px - pis[(i, ch)])
                    if delta > max_change:
                        max_change = delta
        new_pis, pis = pis, new_pis                                                    # :189-190
        new_lambdas, lambdas = lambdas, new_lambdas
        change_per_iter[change_index] = max_change
        change_index = (change_index + 1) % iters_for_convergence
        if max(change_per_iter) <= tol:
            break

    q = int(query)                                                                     # :201-217
    lam = np.ones(card[q]) if not children[q] else _fold_mul(lambdas[(ch, q)] for ch in children[q])
    if not parents[q]:
        pi = factors[q]
    else:
        ft = _broadcast_mul(factors[q], range(1, 1 + len(parents[q])), [pis[(p, q)] for p in parents[q]])
        pi = _sum_axes(ft, range(1, 1 + len(parents[q])))
    out = lam * pi
    _factor_normalize_inplace(out)
    return (out, iters_done) if return_iters else out
