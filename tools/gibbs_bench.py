#!/usr/bin/env python
"""BASELINE config C4: GibbsSamplingNodewise on a 200-node random net, 20 evidence nodes, burn-in 1000, thin 5,
nsamples 2000 (node updates per chain), n_chains chains -> node-updates/s on one GPU, checked against the exact
posterior (4 sigma from between-chain variance) and with the oracle timed beside it on a bounded number of chains.
    python tools/gibbs_bench.py [--chains 65536] [--json out.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def build_case(bn, n_nodes=200, net_seed=0, n_ev=20):
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=net_seed)
    lay = bn.flatten(net)
    rng = np.random.default_rng(4000 + net_seed)
    while True:
        picks = [str(x) for x in rng.choice(lay.names, size=1 + n_ev, replace=False)]
        query, evn = picks[:1], picks[1:]
        evidence = {e: int(rng.integers(1, net.ncategories(e) + 1)) for e in evn}
        im = bn.LikelihoodWeightingInference(nsamples=4096, seed=1)
        im.raw_counts(bn.InferenceState(net, query, evidence))
        if im.last_stats[0] > 0:
            return net, lay, query, evidence


def run(bn, torch, n_chains=65536, nsamples=2000, burn_in=1000, thin=5, reps=5, cpu_chains=256):
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import device_net
    net, lay, query, evidence = build_case(bn)
    bn.set_gibbs_mode(os.environ.get("BN_GIBBS_KERNEL", "specialized"))  # auto would JIT only after 2e11 updates
    dn = device_net(net, 0)
    ev, q = lay.evidence_vector(evidence), lay.query_indices(query)
    ncell = int(np.prod([lay.card[i] for i in q]))
    out = torch.zeros(ncell, dtype=torch.float64, device="cuda:0")
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    L = _lib.lib()

    def launch(seed):
        _lib.check(L.bn_gibbs_infer_dev(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p),
                                        C.c_int32(q.size), C.c_uint64(0), C.c_uint64(n_chains), C.c_int64(nsamples),
                                        C.c_int64(burn_in), C.c_int64(thin), C.c_int32(0), C.c_uint64(seed),
                                        C.c_void_p(out.data_ptr())))
    for w in range(3):
        launch(100 + w)
    torch.cuda.synchronize()
    ms = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(r)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    counts = out.cpu().numpy()
    updates = n_chains * nsamples  # one iteration = one node update (src/Inference/gibbs.jl:120)
    res = {"n_chains": n_chains, "nsamples": nsamples, "burn_in": burn_in, "thin": thin, "n_nodes": int(lay.n),
           "n_evidence": len(evidence), "ms": float(np.median(ms)), "node_updates_per_s": updates / (np.median(ms) * 1e-3),
           "posterior": [round(float(x), 5) for x in counts / counts.sum()]}
    # cross-check against likelihood weighting with 2e9 particles on the same case (variable elimination on this
    # 200-node net needs a 4e10-entry intermediate under min-degree ordering, so there is no exact answer to hold)
    lw = bn.LikelihoodWeightingInference(nsamples=2_000_000_000, seed=7)
    c = lw.raw_counts(bn.InferenceState(net, query, evidence))
    p = c / c.sum()
    res["lw_2e9"] = [round(float(x), 5) for x in p]
    res["lw_n_eff"] = lw.last_stats[0] ** 2 / lw.last_stats[1]
    kept = n_chains * ((nsamples - burn_in + thin - 1) // thin)
    res["max_abs_err"] = float(np.abs(counts / counts.sum() - p).max())
    res["kept_samples"] = int(kept)
    # CPU baseline: the oracle's single-chain loop (gibbs.jl:66-145 restated), OpenMP over chains, bounded sample
    from oracle import capi
    from oracle.factors_np import FlatNet
    on = FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))
    t0 = time.perf_counter()
    capi.gibbs_infer(on, ev, q, cpu_chains, nsamples, burn_in, thin, seed=0)
    dt = time.perf_counter() - t0
    cpu_chains2 = int(max(cpu_chains, min(65536, cpu_chains / dt * 8)))
    t0 = time.perf_counter()
    capi.gibbs_infer(on, ev, q, cpu_chains2, nsamples, burn_in, thin, seed=0)
    dt = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": cpu_chains2 * nsamples / dt, "unit": "node-updates/s", "cores": capi.num_threads(),
                           "kind": "port", "sample": f"{cpu_chains2} chains of the same case"}
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--nsamples", type=int, default=2000)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import torch
    import bayesnets_jl_b200 as bn
    r = run(bn, torch, n_chains=a.chains, nsamples=a.nsamples)
    print(json.dumps(r))
    if a.json:
        json.dump(r, open(a.json, "w"), indent=1)
