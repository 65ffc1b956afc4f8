#!/usr/bin/env python
"""Fuzz bn_factor_permutedims (tiled transpose + contraction fallback) against numpy.transpose on random shapes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bayesnets_jl_b200 as bn

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0 = time.time()
bad = 0
for case in range(n_cases):
    kind = case % 5
    if kind == 0:
        card = [2] * int(rng.integers(1, 19))
    elif kind == 1:
        card = [int(rng.integers(1, 8)) for _ in range(int(rng.integers(1, 9)))]
    elif kind == 2:
        card = [int(rng.choice([1, 2, 3, 4, 16, 31, 32, 33, 64, 100, 255])) for _ in range(int(rng.integers(1, 4)))]
    elif kind == 3:
        card = [int(rng.choice([2, 4, 8])) for _ in range(int(rng.integers(2, 9)))]
    else:
        card = [int(rng.integers(1, 4)) for _ in range(int(rng.integers(8, 15)))]
    while int(np.prod(card)) > 1 << 21:
        card.pop()
    nd = len(card)
    a = rng.random(card)
    f = bn.Factor([f"z{case}_{i}" for i in range(nd)], a)
    for _ in range(3):
        perm = [int(x) for x in rng.permutation(nd)]
        g = f.permutedims([p + 1 for p in perm])
        if not np.array_equal(g.potential, np.transpose(a, perm)):
            bad += 1
            print("MISMATCH", card, perm, flush=True)
print(f"{n_cases} shapes x 3 permutations, mismatches: {bad}, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
