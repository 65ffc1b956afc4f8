"""Host-side model containers: CategoricalCPD / DiscreteCPD and (Discrete)BayesNet.

Mirrors, for the tabular-discrete case only, src/CPDs/categorical_cpd.jl:21-56,105-107,
src/bayes_nets.jl:26-132,287-317 and src/DiscreteBayesNet/discrete_bayes_net.jl:11-47,96-98,
src/gen_bayes_nets.jl.  Node names are strings where the reference uses Symbols; states are 1-based
integers exactly as in the reference (the 0-based shift happens in device_layout.py).
Nothing here is on the hot path: it is the structure the device layout is flattened from.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import math

import numpy as np

Assignment = dict  # Dict{Symbol,Any}: src/ProbabilisticGraphicalModels/assignments.jl:1


def consistent(a: dict, b: dict) -> bool:
    """src/ProbabilisticGraphicalModels/assignments.jl:13-22."""
    return all(not (k in b) or b[k] == v for k, v in a.items())


@dataclass
class CategoricalCPD:
    """P(target | parents); `distributions[q]` is the probability vector for the q-th parental
    instantiation in DMU order -- first parent fastest (src/CPDs/categorical_cpd.jl:1-28)."""
    target: str
    parents: List[str] = field(default_factory=list)
    parental_ncategories: List[int] = field(default_factory=list)
    distributions: List[np.ndarray] = field(default_factory=list)

    def __post_init__(self):
        self.parents = list(self.parents)
        self.parental_ncategories = [int(x) for x in self.parental_ncategories]
        self.distributions = [np.asarray(d, np.float64) for d in self.distributions]
        q = int(np.prod(self.parental_ncategories)) if self.parents else 1
        if len(self.distributions) != q:
            raise ValueError(f"CPD {self.target}: {len(self.distributions)} distributions for {q} parental instantiations")
        k = len(self.distributions[0])
        if any(len(d) != k for d in self.distributions):
            raise ValueError(f"CPD {self.target}: distributions differ in number of categories")

    @property
    def name(self) -> str:
        return self.target

    def ncategories(self) -> int:  # :55-56
        return len(self.distributions[0])

    def row_index(self, a: dict) -> int:
        """LinearIndices(parental_ncategories)[a[p1], a[p2], ...] - 1 (:36-46)."""
        q, stride = 0, 1
        for p, c in zip(self.parents, self.parental_ncategories):
            v = int(a[p])
            if not 1 <= v <= c:
                raise IndexError(f"BoundsError: parent {p} = {v} outside 1:{c}")
            q += stride * (v - 1)
            stride *= c
        return q

    def __call__(self, a: dict | None = None) -> np.ndarray:
        """cpd(a): the conditional distribution (probability vector) given the parents in `a`."""
        return self.distributions[self.row_index(a or {})]

    def pdf(self, a: dict) -> float:  # src/CPDs/cpds.jl:104
        d = self(a)
        x = a[self.target]
        return float(d[x - 1]) if 1 <= x <= len(d) else 0.0


def DiscreteCPD(target, *args) -> CategoricalCPD:
    """DiscreteCPD(target, prob) normalises `prob` (src/CPDs/categorical_cpd.jl:107);
    DiscreteCPD(target, parents, parental_ncategories, distributions) is the plain constructor."""
    if len(args) == 1:
        p = np.asarray(args[0], np.float64)
        return CategoricalCPD(target, [], [], [p / p.sum()])
    parents, pn, dists = args
    return CategoricalCPD(target, parents, pn, dists)


class BayesNet:
    """BayesNet{CategoricalCPD}: DAG + CPDs kept in topological order (src/bayes_nets.jl:50-71)."""

    def __init__(self, cpds: Sequence[CategoricalCPD] = ()):
        self.cpds: List[CategoricalCPD] = []
        self.name_to_index: Dict[str, int] = {}
        for c in cpds:
            self._append(c)
        if self.cpds:
            self._enforce_topological_order()

    # -- construction -----------------------------------------------------------------------
    def _append(self, cpd: CategoricalCPD) -> None:
        if cpd.name in self.name_to_index:
            raise ValueError(f"A CPD with name {cpd.name} already exists!")
        self.cpds.append(cpd)
        self.name_to_index[cpd.name] = len(self.cpds) - 1

    def _enforce_topological_order(self) -> None:
        """topological_sort_by_dfs as Graphs.jl does it: iterative DFS over vertices in index order,
        out-neighbours ascending, reverse post-order (src/bayes_nets.jl:26-48)."""
        n = len(self.cpds)
        for c in self.cpds:
            for p in c.parents:
                if p not in self.name_to_index:
                    raise KeyError(f"parent {p} of {c.name} is not in the network")
        out = [[] for _ in range(n)]
        for j, c in enumerate(self.cpds):
            for p in c.parents:
                out[self.name_to_index[p]].append(j)
        for lst in out:
            lst.sort()
        color = [0] * n
        verts = []
        for v in range(n):
            if color[v]:
                continue
            stack = [v]
            color[v] = 1
            while stack:
                u = stack[-1]
                w = -1
                for nb in out[u]:
                    if color[nb] == 1:
                        raise ValueError("BayesNet graph is non-acyclic!")
                    if color[nb] == 0:
                        w = nb
                        break
                if w >= 0:
                    color[w] = 1
                    stack.append(w)
                else:
                    color[u] = 2
                    verts.append(u)
                    stack.pop()
        order = verts[::-1]
        self.cpds = [self.cpds[j] for j in order]
        self.name_to_index = {c.name: i for i, c in enumerate(self.cpds)}

    def push(self, cpd: CategoricalCPD) -> "BayesNet":
        """push!(bn, cpd): src/bayes_nets.jl:258-276."""
        old = (list(self.cpds), dict(self.name_to_index))
        self._append(cpd)
        try:
            self._enforce_topological_order()
        except Exception:
            self.cpds, self.name_to_index = old
            raise
        return self

    # -- queries ----------------------------------------------------------------------------
    def __len__(self):
        return len(self.cpds)

    def names(self) -> List[str]:  # :80-86
        return [c.name for c in self.cpds]

    def get(self, name) -> CategoricalCPD:
        return self.cpds[name] if isinstance(name, int) else self.cpds[self.name_to_index[name]]

    def parents(self, name: str) -> List[str]:
        return list(self.get(name).parents)

    def children(self, name: str) -> List[str]:  # :96-99, ascending index
        return [c.name for c in self.cpds if name in c.parents]

    def markov_blanket(self, name: str) -> set:  # :124-132
        mb = set()
        for ch in self.children(name):
            mb.update(self.parents(ch))
            mb.add(ch)
        mb.update(self.parents(name))
        mb.discard(name)
        return mb

    def ncategories(self, name: str) -> int:  # discrete_bayes_net.jl:96-98
        return self.get(name).ncategories()

    def pdf(self, assignment: dict | None = None, /, **kw) -> float:  # :311-317
        a = dict(assignment or {}, **kw)
        r = 1.0
        for c in self.cpds:
            r *= c.pdf(a)
        return r

    def logpdf(self, assignment: dict | None = None, /, **kw) -> float:  # :322-328
        """logpdf(bn, assignment): one row, evaluated where the assignment lives (the host), like the reference.
        A whole data table goes through `logpdf(bn, df)` (learning.py), i.e. the counting kernel."""
        a = dict(assignment or {}, **kw)
        r = 0.0
        for c in self.cpds:
            p = c.pdf(a)
            r += math.log(p) if p > 0 else (-math.inf if p == 0 else math.nan)
        return r


DiscreteBayesNet = BayesNet


# ---------------------------------------------------------------------------------------------
# generators (src/gen_bayes_nets.jl, src/DiscreteBayesNet/discrete_bayes_net.jl:31-47)
# ---------------------------------------------------------------------------------------------
def rand_cpd(bn: BayesNet, ncategories: int, target: str, parents: Sequence[str] = (),
             uniform_dirichlet_prior: float = 1.0, rng: np.random.Generator | None = None) -> CategoricalCPD:
    """Random CPT, every column ~ Dirichlet(alpha,...,alpha) (discrete_bayes_net.jl:31-47)."""
    rng = rng or np.random.default_rng()
    if target in bn.name_to_index:
        raise ValueError(f"A CPD with name {target} already exists!")
    pn = [bn.ncategories(p) for p in parents]
    q = int(np.prod(pn)) if parents else 1
    dists = [rng.dirichlet(np.full(ncategories, uniform_dirichlet_prior)) for _ in range(q)]
    return CategoricalCPD(target, list(parents), pn, dists)


def rand_discrete_bn(num_nodes: int = 16, max_num_parents: int = 3, max_num_states: int = 5,
                     connected: bool = True, seed: int | None = None) -> BayesNet:
    """src/gen_bayes_nets.jl:21-61: node i gets rand(2:max_states) states and
    min(i-1, max(min_parents, rand(0:min(i-1, max_parents)))) parents drawn without replacement."""
    if num_nodes <= 0:
        raise ValueError("`num_nodes` must be greater than 0")
    if max_num_parents <= 0:
        raise ValueError("`max_num_parents` must be greater than 0")
    if max_num_states <= 1:
        raise ValueError("`max_num_states` must be greater than 1")
    rng = np.random.default_rng(seed)
    min_parents = 1 if connected else 0
    bn = BayesNet()
    for i in range(1, num_nodes + 1):
        s = f"N{i}"
        n_states = int(rng.integers(2, max_num_states + 1))
        n_par = int(rng.integers(0, min(len(bn), max_num_parents) + 1))
        k = min(len(bn), max(min_parents, n_par))
        parents = [str(x) for x in rng.choice(bn.names(), size=k, replace=False)] if k else []
        bn.push(rand_cpd(bn, n_states, s, parents, rng=rng))
    return bn


def rand_bn_inference(bn: BayesNet, num_query: int = 2, num_evidence: int = 3, seed: int | None = None):
    """src/gen_bayes_nets.jl:69-87 -> (query names, evidence Assignment with 1-based states)."""
    if num_query + num_evidence > len(bn):
        raise ValueError(f"Number of qurey and evidence nodes ({num_query + num_evidence}) is greater than "
                         f"number of nodes ({len(bn)})")
    rng = np.random.default_rng(seed)
    non_hidden = [str(x) for x in rng.choice(bn.names(), size=num_query + num_evidence, replace=False)]
    query = non_hidden[:num_query]
    evidence = {ev: int(rng.integers(1, bn.ncategories(ev) + 1)) for ev in non_hidden[num_query:]}
    return query, evidence


def get_asia_bn() -> BayesNet:
    """src/gen_bayes_nets.jl:124-158 (the ergodic Asia variant, E removed)."""
    bn = BayesNet()
    bn.push(CategoricalCPD("A", [], [], [[0.99, 0.01]]))
    bn.push(CategoricalCPD("S", [], [], [[0.5, 0.5]]))
    bn.push(CategoricalCPD("T", ["A"], [2], [[0.99, 0.01], [0.95, 0.05]]))
    bn.push(CategoricalCPD("L", ["S"], [2], [[0.99, 0.01], [0.9, 0.1]]))
    bn.push(CategoricalCPD("B", ["S"], [2], [[0.97, 0.03], [0.4, 0.6]]))
    bn.push(CategoricalCPD("X", ["T", "L"], [2, 2], [[0.95, 0.05], [0.02, 0.98], [0.02, 0.98], [0.02, 0.98]]))
    bn.push(CategoricalCPD("D", ["T", "L", "B"], [2, 2, 2],
                           [[0.9, 0.1], [0.2, 0.8], [0.3, 0.7], [0.1, 0.9],
                            [0.3, 0.7], [0.1, 0.9], [0.3, 0.7], [0.1, 0.9]]))
    return bn


def get_sprinkler_bn(seed: int | None = None) -> BayesNet:  # :92-100
    rng = np.random.default_rng(seed)
    bn = BayesNet()
    bn.push(rand_cpd(bn, 2, "Rain", [], rng=rng))
    bn.push(rand_cpd(bn, 2, "Sprinkler", ["Rain"], rng=rng))
    bn.push(rand_cpd(bn, 2, "WetGrass", ["Sprinkler", "Rain"], rng=rng))
    return bn


def get_sat_fail_bn(seed: int | None = None) -> BayesNet:  # :105-116
    rng = np.random.default_rng(seed)
    bn = BayesNet()
    bn.push(rand_cpd(bn, 2, "B", [], rng=rng))
    bn.push(rand_cpd(bn, 2, "S", [], rng=rng))
    bn.push(rand_cpd(bn, 2, "E", ["B", "S"], rng=rng))
    bn.push(rand_cpd(bn, 2, "D", ["E"], rng=rng))
    bn.push(rand_cpd(bn, 2, "C", ["E"], rng=rng))
    return bn
