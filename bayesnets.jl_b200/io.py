"""Input formats on either side of the hot path: readxdsl and the text sufficient-statistics format.

Mirrors src/DiscreteBayesNet/io.jl:27-86 (readxdsl; note reverse!(parents), :67, because SMILE varies the
first parent slowest) and :131-228 (write/read of MIME"text/plain").  Host-side parsing only.
"""
from __future__ import annotations

import os
import re
import xml.etree.ElementTree as ET

import numpy as np

from .bayes_net import BayesNet, CategoricalCPD, DiscreteCPD


def readxdsl(filename: str) -> BayesNet:
    if os.path.splitext(filename)[1] != ".xdsl":
        raise ValueError("readxdsl only supports .xdsl format")
    root = ET.parse(filename).getroot()
    nodes = root.find("nodes")
    bn = BayesNet()
    for e in list(nodes):
        name = e.get("id")
        states = []
        for s in e.findall("state"):
            m = re.search(r"\d", s.get("id"))
            assert m is not None, "All state ids must be integers"
            states.append(int(m.group(0)))
        probs = np.array([float(x) for x in e.find("probabilities").text.split()], np.float64)
        k = len(states)
        pe = e.find("parents")
        if pe is not None and pe.text and pe.text.split():
            parents = pe.text.split()[::-1]  # reverse!(parents), io.jl:67
            pn = [bn.ncategories(p) for p in parents]
            q = int(np.prod(pn))
            dists = [probs[k * i:k * (i + 1)] for i in range(q)]
            bn.push(CategoricalCPD(name, parents, pn, dists))
        else:
            bn.push(DiscreteCPD(name, probs))
    return bn


def _fmt(x: float) -> str:
    return "%.16g" % x


def write_text(io, bn: BayesNet) -> None:
    """write(io, MIME"text/plain"(), bn): io.jl:131-168."""
    names = bn.names()
    io.write(" ".join(names) + "\n")
    for i in names:
        io.write("".join("1" if i in bn.parents(j) else "0" for j in names) + "\n")
    io.write(" ".join(str(bn.ncategories(n)) for n in names) + "\n")
    vals = []
    for n in names:
        for d in bn.get(n).distributions:
            vals += [_fmt(float(p)) for p in d[:-1]]
    io.write(" ".join(vals) + "\n")


def read_text(io) -> BayesNet:
    """read(io, MIME"text/plain"(), DiscreteBayesNet): io.jl:170-228."""
    names = io.readline().split()
    n = len(names)
    adj = np.zeros((n, n), bool)
    for i in range(n):
        line = io.readline()
        for j in range(n):
            adj[i, j] = line[j] == "1"
    rs = [int(s) for s in io.readline().split()]
    stats = io.readline().split()
    idx = [0]

    def pull(r):
        p = np.zeros(r)
        for j in range(r - 1):
            p[j] = float(stats[idx[0]])
            idx[0] += 1
        p[-1] = 1.0 - p[:-1].sum()
        if p[-1] < 0.0:
            p[-1] = 0.0
            p /= p.sum()
        return p

    cpds = []
    for i in range(n):
        parents = [j for j in range(n) if adj[j, i]]
        if not parents:
            cpds.append(CategoricalCPD(names[i], [], [], [pull(rs[i])]))
        else:
            pn = [rs[j] for j in parents]
            q = int(np.prod(pn))
            cpds.append(CategoricalCPD(names[i], [names[j] for j in parents], pn, [pull(rs[i]) for _ in range(q)]))
    return BayesNet(cpds)
