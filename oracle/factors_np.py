"""numpy restatement of src/Factors + ExactInference.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Arrays are Fortran-ordered (Julia column-major, src/Factors/factors_main.jl:20); dims are hashable names.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np


class DimensionMismatch(ValueError):
    pass


@dataclass
class NpFactor:
    """Factor(dimensions, potential): src/Factors/factors_main.jl:18-36."""
    dimensions: list
    potential: np.ndarray

    def __post_init__(self):
        self.dimensions = list(self.dimensions)
        arr = np.asarray(self.potential, np.float64)
        self.potential = np.asfortranarray(arr) if arr.ndim else arr
        if len(set(self.dimensions)) != len(self.dimensions):
            raise ValueError("Dimensions must be unique")  # errors.jl:10
        if len(self.dimensions) != self.potential.ndim:
            raise DimensionMismatch("`potential` must have as many dimensions as dims")

    @property
    def size(self):
        return self.potential.shape

    def __len__(self):
        return int(self.potential.size)

    # -- join(*, phi1, phi2, :outer): src/Factors/factors_dims.jl:185-234,269
    def __mul__(self, other: "NpFactor") -> "NpFactor":
        p1, p2 = self, other
        if len(p1) < len(p2):  # :189-191
            p1, p2 = p2, p1
        common = [d for d in p1.dimensions if d in p2.dimensions]
        ic1 = [p1.dimensions.index(d) for d in common]
        ic2 = [p2.dimensions.index(d) for d in common]
        if [p1.size[i] for i in ic1] != [p2.size[i] for i in ic2]:
            raise DimensionMismatch("Common dimensions must have same size")  # :197-199
        new_dims = p1.dimensions + [d for d in p2.dimensions if d not in common]  # :203
        if p2.potential.ndim != 0:
            unique2 = [d for d in p2.dimensions if d not in common]
            iu2 = [p2.dimensions.index(d) for d in unique2]
            temp = np.transpose(p2.potential, ic2 + iu2)  # permutedims(phi2, [common..., unique2...]) :213
            # reshape_lengths (:216-222): singleton axes wherever phi1 has a dim phi2 lacks.
            # `common` was collected in phi1 order, so ic1 is ascending and only axes are inserted.
            index = tuple(slice(None) if i in ic1 else None for i in range(p1.potential.ndim))
            temp = temp[index + (slice(None),) * len(iu2)]
            a = p1.potential.reshape(p1.size + (1,) * len(iu2), order="F")
            new_v = a * temp  # broadcast!(*, new_v, phi1.potential, temp) :233
        else:
            new_v = p1.potential * p2.potential.reshape(())
        return NpFactor(new_dims, new_v)

    # -- join(op, phi1, phi2, :outer) for / + -: src/Factors/factors_dims.jl:185-234,270-272.  The swap to "larger
    # factor first" (:189-191) happens BEFORE op is applied, so / and - depend on the operand sizes, as in the reference.
    def join(self, op: str, other: "NpFactor") -> "NpFactor":
        f = {"*": np.multiply, "/": np.divide, "+": np.add, "-": np.subtract}[op]
        p1, p2 = self, other
        if len(p1) < len(p2):
            p1, p2 = p2, p1
        common = [d for d in p1.dimensions if d in p2.dimensions]
        if [p1.size[p1.dimensions.index(d)] for d in common] != [p2.size[p2.dimensions.index(d)] for d in common]:
            raise DimensionMismatch("Common dimensions must have same size")
        unique2 = [d for d in p2.dimensions if d not in common]
        new_dims = p1.dimensions + unique2
        if p2.potential.ndim == 0:
            with np.errstate(all="ignore"):
                return NpFactor(new_dims, f(p1.potential, p2.potential.reshape(())))
        temp = np.transpose(p2.potential, [p2.dimensions.index(d) for d in common + unique2])
        index = tuple(slice(None) if d in common else None for d in p1.dimensions)
        temp = temp[index + (slice(None),) * len(unique2)]
        a = p1.potential.reshape(p1.size + (1,) * len(unique2), order="F")
        with np.errstate(all="ignore"):
            return NpFactor(new_dims, f(a, temp))

    # -- reducedim(op, phi, dims) for + * max min: src/Factors/factors_dims.jl:60-85,100-107
    def reducedim(self, op: str, dims, pre=None) -> "NpFactor":
        f = {"+": np.add, "*": np.multiply, "max": np.maximum, "min": np.minimum}[op]  # np.maximum propagates NaN like Base.max
        dims = [dims] if not isinstance(dims, (list, tuple)) else list(dims)
        for d in dims:
            if d not in self.dimensions:
                raise ValueError(f"{d} is not a valid dimension")
        axes = sorted(self.dimensions.index(d) for d in dims)
        keep = [d for d in self.dimensions if d not in dims]
        src = self.potential if pre is None else pre(self.potential)
        if not axes:
            return NpFactor(keep, src.copy())
        moved = np.moveaxis(src, axes, range(len(axes)))
        flat = moved.reshape((-1,) + moved.shape[len(axes):], order="F")
        acc = flat[0].copy()
        for t in range(1, flat.shape[0]):
            acc = f(acc, flat[t])
        return NpFactor(keep, acc)

    # -- normalize!(phi, dims; p): src/Factors/factors_dims.jl:24-43 (p = 2 divides by sum(abs2): no square root)
    def normalize_dims(self, dims, p: int = 1) -> "NpFactor":
        dims = list(dict.fromkeys([dims] if not isinstance(dims, (list, tuple)) else dims))
        if p not in (1, 2):
            raise ValueError(f"p = {p} is not supported")
        if not dims:
            return NpFactor(self.dimensions, self.potential.copy())
        total = self.reducedim("+", dims, pre=np.abs if p == 1 else np.square)
        index = tuple(None if d in dims else slice(None) for d in self.dimensions)
        with np.errstate(all="ignore"):
            return NpFactor(self.dimensions, self.potential / total.potential[index])

    # -- broadcast(op, phi, dims, values): src/Factors/factors_dims.jl:118-165; several values fold left to right
    def broadcast(self, op: str, dims, values) -> "NpFactor":
        f = {"*": np.multiply, "/": np.divide, "+": np.add, "-": np.subtract, "max": np.maximum, "min": np.minimum}[op]
        if not isinstance(dims, (list, tuple)):
            dims, values = [dims], [values]
        if len(set(dims)) != len(dims):
            raise ValueError("Dimensions must be unique")
        for d in dims:
            if d not in self.dimensions:
                raise ValueError(f"{d} is not a valid dimension")
        if len(dims) != len(values):
            raise ValueError("Number of dimensions does not match number of values to broadcast")
        out = self.potential.copy(order="F")
        for d, v in zip(dims, values):
            v = np.atleast_1d(np.asarray(v, np.float64))
            ax = self.dimensions.index(d)
            if v.size not in (1, self.size[ax]):
                raise DimensionMismatch(f"broadcast value for {d} has length {v.size}")
            shape = [1] * out.ndim
            shape[ax] = v.size
            with np.errstate(all="ignore"):
                out = f(out, v.reshape(shape))
        return NpFactor(self.dimensions, out)

    # -- reducedim(+, phi, dims): src/Factors/factors_dims.jl:60-85,100
    def sum(self, dims) -> "NpFactor":
        dims = [dims] if not isinstance(dims, (list, tuple)) else list(dims)
        for d in dims:
            if d not in self.dimensions:
                raise ValueError(f"{d} is not a valid dimension")  # errors.jl:12
        axes = tuple(self.dimensions.index(d) for d in dims)
        keep = [d for d in self.dimensions if d not in dims]
        if not axes:
            return NpFactor(keep, self.potential.copy())
        # sequential accumulation along the reduced axes in column-major order
        # summed axes first, ordered by their position in the factor (column-major sweep order)
        order = np.argsort(axes)
        moved = np.moveaxis(self.potential, [axes[o] for o in order], range(len(axes)))
        flat = moved.reshape((-1,) + moved.shape[len(axes):], order="F")
        acc = flat[0].copy()
        for t in range(1, flat.shape[0]):
            acc = acc + flat[t]
        return NpFactor(keep, acc)

    # -- normalize!(phi; p): src/Factors/factors_dims.jl:45-57
    def normalize(self, p: int = 1) -> "NpFactor":
        if p == 1:
            total = np.abs(self.potential).sum()
        elif p == 2:
            total = (self.potential ** 2).sum()
        else:
            raise ValueError(f"p = {p} is not supported")
        with np.errstate(invalid="ignore", divide="ignore"):
            return NpFactor(self.dimensions, self.potential / total)

    # -- getindex(phi, Assignment): src/Factors/factors_access.jl:19-35,48-72 (0-based states here)
    def slice(self, assignment: dict) -> "NpFactor":
        idx = []
        keep = []
        for i, d in enumerate(self.dimensions):
            if d in assignment:
                v = assignment[d]
                if not isinstance(v, (int, np.integer)):
                    raise TypeError(f"Invalid state for dimension {d}")
                if v < 0 or v >= self.size[i]:
                    raise IndexError(f"BoundsError({d}, {v})")
                idx.append(int(v))
            else:
                idx.append(slice(None))
                keep.append(d)
        return NpFactor(keep, self.potential[tuple(idx)])

    # -- setindex!(phi, v, Assignment): src/Factors/factors_access.jl:39-46 (0-based states here).  Dims of the
    #    assignment that phi lacks are ignored and unassigned dims are Colons (_translate_index, :48-72); with a Colon
    #    the value is broadcast into the slab (`.= v`), otherwise one cell is set.  In place, like the reference.
    def setindex(self, v, assignment: dict) -> "NpFactor":
        idx = []
        for i, d in enumerate(self.dimensions):
            if d in assignment:
                a = assignment[d]
                if not isinstance(a, (int, np.integer)):
                    raise TypeError(f"Invalid state for dimension {d}")
                if a < 0 or a >= self.size[i]:
                    raise IndexError(f"BoundsError({d}, {a})")
                idx.append(int(a))
            else:
                idx.append(slice(None))
        self.potential[tuple(idx)] = v
        return self

    # -- permutedims(phi, perm): src/Factors/factors_main.jl:160-166 (perm = 0-based positions here)
    def permutedims(self, perm) -> "NpFactor":
        perm = [int(x) for x in perm]
        if sorted(perm) != list(range(len(self.dimensions))):
            raise ValueError("no valid permutation of dimensions")  # Base.permutedims: ArgumentError
        return NpFactor([self.dimensions[i] for i in perm], np.asfortranarray(np.transpose(self.potential, perm)))

    # -- push!(phi, dim, length): src/Factors/factors_main.jl:148-158 over duplicate, auxiliary.jl:50-65: a new LAST
    #    dimension along which the whole table is repeated
    def push(self, dim, length: int) -> "NpFactor":
        if dim in self.dimensions:
            raise RuntimeError(f"Dimension {dim} already exists")
        pot = np.asfortranarray(np.repeat(self.potential[..., None], int(length), axis=-1))
        return NpFactor(list(self.dimensions) + [dim], pot)

    def permuted_to(self, dims) -> np.ndarray:
        """Potential with axes reordered to `dims` (for order-independent comparison)."""
        return np.transpose(self.potential, [self.dimensions.index(d) for d in dims])


# --------------------------------------------------------------------------------------------
# flat nets for the oracle's own use
# --------------------------------------------------------------------------------------------
@dataclass
class FlatNet:
    names: list
    card: np.ndarray
    par_ptr: np.ndarray
    par_idx: np.ndarray
    cpt_off: np.ndarray
    cpt: np.ndarray
    index: dict = field(default_factory=dict)

    def parents(self, i):
        return [int(p) for p in self.par_idx[self.par_ptr[i]:self.par_ptr[i + 1]]]

    def children(self, i):
        return [j for j in range(len(self.names)) if i in self.parents(j)]


def make_net(nodes) -> FlatNet:
    """nodes: [(name, parents(list of names), rows)] in topological order; rows = list of probability
    vectors in DMU order (first parent fastest), i.e. CategoricalCPD.distributions
    (src/CPDs/categorical_cpd.jl:21-28)."""
    names = [n for n, _, _ in nodes]
    index = {n: i for i, n in enumerate(names)}
    card, par_ptr, par_idx, cpt_off, cpt = [], [0], [], [0], []
    for name, parents, rows in nodes:
        rows = [np.asarray(r, np.float64) for r in rows]
        k = len(rows[0])
        for p in parents:
            assert index[p] < index[name], "nodes must be topologically ordered"
        q = int(np.prod([card[index[p]] for p in parents])) if parents else 1
        assert len(rows) == q, (name, len(rows), q)
        assert all(len(r) == k for r in rows)
        card.append(k)
        par_idx += [index[p] for p in parents]
        par_ptr.append(len(par_idx))
        cpt += [x for r in rows for x in r]
        cpt_off.append(len(cpt))
    return FlatNet(names, np.array(card, np.int32), np.array(par_ptr, np.int32), np.array(par_idx, np.int32),
                   np.array(cpt_off, np.int64), np.array(cpt, np.float64), index)


def cpd_factor(net: FlatNet, i: int) -> NpFactor:
    """convert(Factor, cpd): dims = [node, parents...], column-major copy (src/Factors/factors_main.jl:62-68)."""
    ps = net.parents(i)
    shape = [int(net.card[i])] + [int(net.card[p]) for p in ps]
    data = net.cpt[net.cpt_off[i]:net.cpt_off[i + 1]].reshape(shape, order="F")
    return NpFactor([net.names[i]] + [net.names[p] for p in ps], data)


def elimination_order(net: FlatNet, query, evidence: dict):
    """Greedy minimum-degree order on the moralised non-evidence graph, query never eliminated.
    The reference delegates to CliqueTrees.jl MMD with the query clique last (src/Inference/exact.jl:37-84);
    that order is not pinned by any test (only posteriors are), and it only changes fp64 rounding."""
    n = len(net.names)
    alive = [i for i in range(n) if net.names[i] not in evidence]
    adj = {i: set() for i in alive}
    for j in range(n):
        fam = [p for p in net.parents(j) if p in adj] + ([j] if j in adj else [])
        for a, b in itertools.combinations(fam, 2):
            adj[a].add(b); adj[b].add(a)
    q = {net.index[x] for x in query}
    order = []
    todo = [i for i in alive if i not in q]
    while todo:
        v = min(todo, key=lambda i: (len(adj[i]), i))
        nb = list(adj[v])
        for a, b in itertools.combinations(nb, 2):
            adj[a].add(b); adj[b].add(a)
        for a in nb:
            adj[a].discard(v)
        del adj[v]
        todo.remove(v)
        order.append(net.names[v])
    return order


def exact_infer(net: FlatNet, query, evidence: dict | None = None) -> NpFactor:
    """infer(::ExactInference): src/Inference/exact.jl:6-33.  evidence maps name -> 0-based state."""
    evidence = evidence or {}
    query = [query] if not isinstance(query, (list, tuple)) else list(dict.fromkeys(query))
    for qn in query:  # InferenceState validation, src/ProbabilisticGraphicalModels/inference.jl:6-14
        if qn not in net.index:
            raise ValueError(f"Query {qn} is not in the probabilistic graphical model")
        if qn in evidence:
            raise ValueError(f"Query {qn} is part of the evidence")
    hidden = elimination_order(net, query, evidence)
    factors = [cpd_factor(net, i).slice(evidence) for i in range(len(net.names))]  # :16
    for h in hidden:  # :20-29
        contain = [f for f in factors if h in f.dimensions]
        if contain:
            factors = [f for f in factors if h not in f.dimensions]
            prod = contain[0]
            for f in contain[1:]:
                prod = prod * f
            factors.append(prod.sum(h))
    phi = factors[0]
    for f in factors[1:]:
        phi = phi * f
    return phi.normalize()  # :31


def brute_force_posterior(net: FlatNet, query, evidence: dict | None = None) -> np.ndarray:
    """Joint enumeration (pdf(bn, a) = prod pdf(cpd, a), src/bayes_nets.jl:311-317); small nets only.
    Returns P(query | evidence) with axes in `query` order."""
    evidence = evidence or {}
    query = [query] if not isinstance(query, (list, tuple)) else list(query)
    n = len(net.names)
    cards = [int(c) for c in net.card]
    qi = [net.index[x] for x in query]
    out = np.zeros([cards[i] for i in qi])
    ranges = [[evidence[net.names[i]]] if net.names[i] in evidence else range(cards[i]) for i in range(n)]
    for x in itertools.product(*ranges):
        p = 1.0
        for i in range(n):
            qrow, stride = 0, 1
            for par in net.parents(i):
                qrow += stride * x[par]; stride *= cards[par]
            p *= net.cpt[net.cpt_off[i] + x[i] + cards[i] * qrow]
        out[tuple(x[i] for i in qi)] += p
    return out / out.sum()
