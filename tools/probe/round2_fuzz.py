#!/usr/bin/env python
"""Differential fuzzing of the round-2 kernels on the GPU (no oracle needed: two device paths must agree bit for bit).

  sums : sum(phi, dims) through the one-shot streaming kernel vs the contraction kernel (BN_FACTOR_NO_STREAM), random
         mixed-radix shapes and dim sets; also push! through dup_kernel vs the contraction kernel.
  gibbs: chains of the NVRTC Gibbs kernel with tabulated conditionals (random row limits) vs without tables
         (BN_GIBBS_TABLE_ROWS=0) on random nets: counts and final states identical.

    python tools/probe/round2_fuzz.py [--sums 2000] [--nets 40]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)


def fuzz_sums(bn, n_cases, rng):
    bad, stream_hits = [], 0
    for case in range(n_cases):
        nd = int(rng.integers(1, 11))
        cards = [int(rng.choice([1, 2, 2, 2, 3, 4, 5, 6, 8])) for _ in range(nd)]
        while np.prod(cards) > (1 << 18):
            cards[int(rng.integers(0, nd))] = 2
        names = [f"fz{case}_{i}" for i in range(nd)]
        A = rng.random(cards)
        f = bn.Factor(names, A)
        k = int(rng.integers(1, nd + 1))
        sd = sorted(int(x) for x in rng.choice(nd, size=k, replace=False))
        os.environ.pop("BN_FACTOR_NO_STREAM", None)
        n0 = bn.launch_count()
        a = f.sum([names[i] for i in sd]).potential
        os.environ["BN_FACTOR_NO_STREAM"] = "1"
        b = f.sum([names[i] for i in sd]).potential
        os.environ.pop("BN_FACTOR_NO_STREAM", None)
        ref = A.sum(axis=tuple(sd))
        if not (np.array_equal(a, b) and np.allclose(a, ref, rtol=1e-12, atol=0)):
            bad.append({"cards": cards, "sum": sd})
        length = int(rng.integers(1, 5))
        g = bn.Factor(names, A)
        g.push_(f"fz{case}_new", length)
        got = g.potential
        if not all(np.array_equal(got[..., j], A) for j in range(length)):
            bad.append({"cards": cards, "push": length})
    return bad


def fuzz_gibbs(bn, n_nets, rng):
    bad = []
    bn.set_gibbs_mode("specialized")
    for case in range(n_nets):
        n_nodes = int(rng.integers(6, 45))
        net = bn.rand_discrete_bn(n_nodes, int(rng.integers(1, 4)), int(rng.integers(2, 6)), seed=int(rng.integers(0, 1 << 30)))
        lay = bn.flatten(net)
        picks = [str(x) for x in rng.choice(lay.names, size=min(n_nodes - 1, 1 + int(rng.integers(0, 4))), replace=False)]
        query, evn = picks[:1], picks[1:]
        evidence = {e: int(rng.integers(1, net.ncategories(e) + 1)) for e in evn}
        full = bool(rng.integers(0, 2))
        cls = bn.GibbsSamplingFull if full else bn.GibbsSamplingNodewise
        ns, burn, thin = (30, 8, 2) if full else (700, 150, 3)
        res = []
        for rows in ("0", str(int(rng.choice([4, 16, 64, 256, 4096])))):
            os.environ["BN_GIBBS_TABLE_ROWS"] = rows
            im = cls(nsamples=ns, burn_in=burn, thin=thin, n_chains=160, seed=case, keep_chain_counts=True)
            c = im.raw_counts(bn.InferenceState(net, query, evidence))
            res.append((c.copy(), im.chain_states.copy(), im.last_chain_counts.copy()))
        os.environ.pop("BN_GIBBS_TABLE_ROWS", None)
        if not all(np.array_equal(x, y) for x, y in zip(res[0], res[1])):
            bad.append({"case": case, "n_nodes": n_nodes, "full": full, "rows": rows})
    bn.set_gibbs_mode("auto")
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sums", type=int, default=2000)
    ap.add_argument("--nets", type=int, default=40)
    a = ap.parse_args()
    import bayesnets_jl_b200 as bn
    rng = np.random.default_rng(20261020)
    t0 = time.time()
    bad_s = fuzz_sums(bn, a.sums, rng)
    t1 = time.time()
    bad_g = fuzz_gibbs(bn, a.nets, rng)
    print(json.dumps({"sum_and_push_cases": a.sums, "sum_mismatches": bad_s[:5], "n_sum_mismatches": len(bad_s),
                      "gibbs_nets": a.nets, "gibbs_mismatches": bad_g[:5], "n_gibbs_mismatches": len(bad_g),
                      "seconds": [round(t1 - t0, 1), round(time.time() - t1, 1)]}))
    sys.exit(1 if (bad_s or bad_g) else 0)
