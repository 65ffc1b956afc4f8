#!/usr/bin/env python
"""Differential fuzzing of the round-2 kernels on the GPU (no oracle needed: two device paths must agree bit for bit).

  sums : sum(phi, dims) through the one-shot streaming kernel vs the contraction kernel (BN_FACTOR_NO_STREAM), random
         mixed-radix shapes and dim sets; also push! through dup_kernel vs the contraction kernel.
  gibbs: chains of the NVRTC Gibbs kernel with tabulated conditionals (random row limits) vs without tables
         (BN_GIBBS_TABLE_ROWS=0) on random nets: counts and final states identical.

  stats: statistics of a device table through the row-lane kernel vs the node-lane kernel (BN_STATS_NODELANES).

    python tools/probe/round2_fuzz.py [--sums 2000] [--nets 40] [--tables 60]
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


def fuzz_stats(bn, n_cases, rng):
    """statistics of a sampled device table: row-lane kernel vs node-lane kernel (BN_STATS_NODELANES), random nets of
    1..160 nodes (every warp count the plan can choose, odd node counts, ragged last tiles, pitch > rows)."""
    import ctypes as C
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import DeviceNet
    from bayesnets_jl_b200.learning import statistics_device
    bad = []
    L = _lib.lib()
    for case in range(n_cases):
        n_nodes = int(rng.integers(2, 161))
        net = bn.rand_discrete_bn(n_nodes, int(rng.integers(1, min(4, n_nodes))), int(rng.integers(2, 6)), seed=int(rng.integers(0, 1 << 30)))
        lay = bn.flatten(net)
        dn = DeviceNet(lay, 0)
        n_rows = int(rng.choice([1, 31, 255, 256, 257, 1000, 4096, 70001, 262144]))
        pitch = n_rows if rng.random() < 0.5 else (n_rows + 15) // 16 * 16 + 16 * int(rng.integers(0, 3))
        tab = _lib.DeviceBuffer(lay.n * pitch, 0, zero=True)
        ev = np.full(lay.n, -1, np.int8)
        if pitch == n_rows:
            _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), C.c_int32(0), C.c_int32(1), C.c_uint64(0),
                                       C.c_uint64(n_rows), C.c_uint64(case), C.c_void_p(tab.ptr), None, None))
        else:  # padded pitch: upload a host table
            host = np.zeros((lay.n, pitch), np.uint8)
            for i in range(lay.n):
                host[i, :n_rows] = rng.integers(0, lay.card[i], n_rows)
            tab.upload(host)
        res = []
        for nl in (None, "1"):
            if nl:
                os.environ["BN_STATS_NODELANES"] = nl
            else:
                os.environ.pop("BN_STATS_NODELANES", None)
            _, c = statistics_device(dn, tab.ptr, n_rows, pitch)
            res.append(c)
        os.environ.pop("BN_STATS_NODELANES", None)
        ok = np.array_equal(res[0], res[1]) and int(res[0][lay.cpt_off[0]:lay.cpt_off[1]].sum()) == n_rows
        if not ok:
            bad.append({"case": case, "n_nodes": n_nodes, "rows": n_rows, "pitch": pitch})
        tab.close()
        dn.close()
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sums", type=int, default=2000)
    ap.add_argument("--nets", type=int, default=40)
    ap.add_argument("--tables", type=int, default=60)
    a = ap.parse_args()
    import bayesnets_jl_b200 as bn
    rng = np.random.default_rng(20261020)
    t0 = time.time()
    bad_s = fuzz_sums(bn, a.sums, rng)
    t1 = time.time()
    bad_g = fuzz_gibbs(bn, a.nets, rng)
    t2 = time.time()
    bad_t = fuzz_stats(bn, a.tables, rng)
    print(json.dumps({"stats_tables": a.tables, "stats_mismatches": bad_t[:5], "n_stats_mismatches": len(bad_t),
                      "stats_seconds": round(time.time() - t2, 1)}))
    print(json.dumps({"sum_and_push_cases": a.sums, "sum_mismatches": bad_s[:5], "n_sum_mismatches": len(bad_s),
                      "gibbs_nets": a.nets, "gibbs_mismatches": bad_g[:5], "n_gibbs_mismatches": len(bad_g),
                      "seconds": [round(t1 - t0, 1), round(time.time() - t1, 1)]}))
    sys.exit(1 if (bad_s or bad_g or bad_t) else 0)
