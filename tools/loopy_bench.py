#!/usr/bin/env python
"""LoopyBelief batched over evidence sets on the C2 network: `--inst` (query, evidence) pairs, each with
10 random evidence nodes and one random query, `--iters` iterations with tol < 0 (fixed work), one kernel launch.
Reports instance-iterations/s and message updates/s on one GPU, checks a few instances bit for bit against the oracle
(numpy restatement on a few pairs, the C restatement -- also the CPU baseline, OpenMP over pairs -- on a bounded sample).
    python tools/loopy_bench.py [--inst 65536] [--iters 50] [--json out.json]
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


def build_case(bn, n_inst, n_nodes=100, net_seed=0, n_ev=10):
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=net_seed)
    lay = bn.flatten(net)
    rng = np.random.default_rng(6000 + net_seed)
    ev = np.full((n_inst, lay.n), -1, np.int8)
    q = np.zeros(n_inst, np.int32)
    for b in range(n_inst):
        picks = rng.choice(lay.n, size=n_ev + 1, replace=False)
        q[b] = picks[0]
        for i in picks[1:]:
            ev[b, i] = rng.integers(0, lay.card[i])
    return net, lay, ev, q


def run(bn, torch, n_inst=65536, iters=50, reps=5, n_check=3, n_nodes=100):
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import device_net
    net, lay, ev, q = build_case(bn, n_inst, n_nodes=n_nodes)
    dn = device_net(net, 0)
    stride = int(lay.card.max())
    ev_d = torch.from_numpy(ev).to("cuda:0")
    q_d = torch.from_numpy(q).to("cuda:0")
    out_d = torch.zeros((n_inst, stride), dtype=torch.float64, device="cuda:0")
    it_d = torch.zeros(n_inst, dtype=torch.int32, device="cuda:0")
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    L = _lib.lib()

    def launch(n_it, tol):
        _lib.check(L.bn_loopy_infer_dev(_lib.h64(dn.handle), C.c_void_p(ev_d.data_ptr()), C.c_void_p(q_d.data_ptr()),
                                        C.c_uint64(n_inst), C.c_int32(n_it), C.c_double(tol), C.c_int32(6), C.c_int32(stride),
                                        C.c_void_p(out_d.data_ptr()), C.c_void_p(it_d.data_ptr())))
    for _ in range(3):
        launch(iters, -1.0)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(iters, -1.0)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms)) * 1e-3
    n_edges = int(lay.par_ptr[-1])
    with_children = int(len(set(int(p) for p in lay.par_idx)))
    res = {"n_nodes": int(lay.n), "n_edges": n_edges, "n_inst": n_inst, "iters": iters, "ms": t * 1e3,
           "instance_iterations_per_s": n_inst * iters / t,
           "message_updates_per_s": n_inst * iters * (n_edges + with_children) / t,
           "instances_per_s_at_this_iteration_count": n_inst / t}
    post = out_d.cpu().numpy()
    # end to end through the host-buffer entry point: evidence rows + queries up, beliefs + iteration counts down
    out_h = np.zeros((n_inst, stride), np.float64)
    it_h = np.zeros(n_inst, np.int32)

    def host_call():
        _lib.check(L.bn_loopy_infer(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p), C.c_uint64(n_inst),
                                    C.c_int32(iters), C.c_double(-1.0), C.c_int32(6), C.c_int32(stride),
                                    _lib.ptr(out_h, _lib.f64p), _lib.ptr(it_h, _lib.i32p)))
    host_call()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); host_call(); ts.append(time.perf_counter() - t0)
    res["e2e"] = {"value": n_inst * iters / float(np.median(ts)), "unit": "instance-iterations/s",
                  "h2d_bytes_per_call": int(ev.nbytes + q.nbytes), "d2h_bytes_per_call": int(out_h.nbytes + it_h.nbytes),
                  "api": "bn_loopy_infer (host buffers)", "identical_to_device_call": bool(np.array_equal(out_h, post, equal_nan=True))}
    # default stopping rule (tol 1e-8, cap 500): how many iterations the instances actually need
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch(500, 1e-8)
    e1.record()
    torch.cuda.synchronize()
    its = it_d.cpu().numpy()
    res["default_rule"] = {"ms": e0.elapsed_time(e1), "iters_mean": float(its.mean()), "iters_max": int(its.max()),
                           "converged_frac": float((its < 500).mean())}
    launch(iters, -1.0)
    torch.cuda.synchronize()
    # parity: the numpy oracle (reference's aliasing kept literally) on a few pairs, the C oracle on a bounded sample;
    # the C oracle with OpenMP over pairs on all host cores is also the CPU baseline
    from oracle import capi
    from oracle.factors_np import FlatNet
    from oracle.loopy_np import loopy_infer
    on = FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))
    same = True
    for b in range(n_check):
        o = loopy_infer(on, int(q[b]), ev[b], nsamples=iters, tol=-1.0)
        same = same and np.array_equal(post[b, :len(o)], o, equal_nan=True)
    res["identical_to_oracle"] = bool(same)
    n_cpu = min(n_inst, 64)
    t0 = time.perf_counter()
    capi.loopy_infer(on, q[:n_cpu], ev[:n_cpu], nsamples=iters, tol=-1.0)
    dt = time.perf_counter() - t0
    n_cpu = int(min(n_inst, max(n_cpu, n_cpu * 8.0 / max(dt, 1e-3))))
    t0 = time.perf_counter()
    co, _ = capi.loopy_infer(on, q[:n_cpu], ev[:n_cpu], nsamples=iters, tol=-1.0)
    dt = time.perf_counter() - t0
    res["identical_to_c_oracle"] = {"pairs": n_cpu, "identical": bool(np.array_equal(post[:n_cpu], co, equal_nan=True))}
    res["cpu_baseline"] = {"value": n_cpu * iters / dt, "unit": "instance-iterations/s", "cores": capi.num_threads(), "kind": "port",
                           "sample": f"{n_cpu} pairs x {iters} iterations of the same case (C restatement, OpenMP over pairs)"}
    _lib.set_stream(None)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--inst", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--check", type=int, default=3)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import torch
    import bayesnets_jl_b200 as bn
    r = run(bn, torch, a.inst, a.iters, n_check=a.check, n_nodes=a.nodes)
    print(json.dumps(r))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(r, f, indent=1)
