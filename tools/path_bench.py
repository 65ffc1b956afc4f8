#!/usr/bin/env python
"""Every SURVEY.md section-8a row that has a kernel, measured on one B200 with the oracle timed beside it.

    python tools/path_bench.py [--json out.json] [--quick]

Each entry: {"row", "api", "config", "unit", "gpu": value, "cpu": {"value", "cores", "kind": "port", "sample"},
"check": what was compared with the oracle / exact answer}.  GPU timings are CUDA events on the launching stream after
3 warm-ups (median of 5); e2e rows go through the public Python mirror with host buffers.
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
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def ev_time(torch, fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def c5_layout(bn, k=23, n_nodes=30, seed=0):
    """BASELINE C5: 30 binary nodes, the last one with k parents (CPT = 2^(k+1) doubles); the other nodes get <= 2
    parents among earlier nodes.  Built as flat arrays (a Python CPD object per row would be 8M objects)."""
    from bayesnets_jl_b200.device_layout import DeviceLayout
    rng = np.random.default_rng(seed)
    names = [f"B{i + 1}" for i in range(n_nodes)]
    card = np.full(n_nodes, 2, np.int32)
    par_ptr, par_idx, cpt_off, chunks = [0], [], [0], []
    for i in range(n_nodes):
        if i == n_nodes - 1:
            ps = sorted(int(x) for x in rng.choice(i, size=k, replace=False))
        else:
            ps = sorted(int(x) for x in rng.choice(i, size=min(i, int(rng.integers(0, 3))), replace=False)) if i else []
        par_idx += ps
        par_ptr.append(len(par_idx))
        rows = 1 << len(ps)
        p0 = rng.random(rows) * 0.9 + 0.05
        tab = np.empty(2 * rows)
        tab[0::2], tab[1::2] = p0, 1.0 - p0
        chunks.append(tab)
        cpt_off.append(cpt_off[-1] + tab.size)
    return DeviceLayout(names, {n: i for i, n in enumerate(names)}, card, np.asarray(par_ptr, np.int32),
                        np.asarray(par_idx, np.int32), np.asarray(cpt_off, np.int64), np.concatenate(chunks))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    import torch
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import DeviceNet, device_net
    from oracle import capi
    from oracle.factors_np import FlatNet, exact_infer
    import bench
    capi.build()
    capi.set_num_threads(len(os.sched_getaffinity(0)))
    cores = capi.num_threads()
    L = _lib.lib()
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    out = []

    def onet(lay):
        return FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))

    net, lay, query, evidence = bench.build_case(bn)
    # the network-specialised kernels, compiled up front (the "auto" modes JIT only after enough cumulative work)
    bn.set_lw_mode("specialized", True)
    bn.set_gibbs_mode("specialized")
    on = onet(lay)
    dn = device_net(net, 0)
    ev, q = lay.evidence_vector(evidence), lay.query_indices(query)
    no_ev = np.full(lay.n, -1, np.int8)

    # ---- a5 / a4 / a6: table samplers, device-resident output (node-major uint8, one column per node)
    n_rows = 1 << 24 if not a.quick else 1 << 20
    st = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
    w = torch.empty(n_rows, dtype=torch.float64, device="cuda:0")
    ok = torch.empty(n_rows, dtype=torch.uint8, device="cuda:0")
    one_ev = np.full(lay.n, -1, np.int8)
    one_ev[lay.index[next(iter(evidence))]] = ev[lay.index[next(iter(evidence))]]
    for row, name, kind, evv, cite in [
            ("a5", "rand(bn, n) ancestral table", 0, no_ev, "src/sampling.jl:24-41,52-57"),
            ("a4", "get_weighted_dataframe (LW rows)", 2, ev, "src/sampling.jl:102-115,153-174"),
            ("a6", "rand(bn, n, evidence) rejection, 1 evidence var, <=100 tries", 1, one_ev, "src/sampling.jl:72-93")]:
        def fn(seed=[0]):
            seed[0] += 1
            _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(evv, _lib.i8p), C.c_int32(kind), C.c_int32(100),
                                       C.c_uint64(0), C.c_uint64(n_rows), C.c_uint64(seed[0]), C.c_void_p(st.data_ptr()),
                                       C.c_void_p(w.data_ptr()), C.c_void_p(ok.data_ptr())))
        ms = ev_time(torch, fn)
        n_cpu = 200_000
        t0 = time.perf_counter()
        cs, cw, cok = capi.sample(on, n_cpu, evv, kind=kind, seed=3)
        dt = time.perf_counter() - t0
        # parity on a prefix: same seed, same rows
        _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(evv, _lib.i8p), C.c_int32(kind), C.c_int32(100),
                                   C.c_uint64(0), C.c_uint64(n_rows), C.c_uint64(3), C.c_void_p(st.data_ptr()),
                                   C.c_void_p(w.data_ptr()), C.c_void_p(ok.data_ptr())))
        same = bool(np.array_equal(st[:, :n_cpu].cpu().numpy(), cs))
        out.append({"row": row, "api": name, "cite": cite, "config": f"C2 net, {n_rows} rows x {lay.n} nodes (uint8)",
                    "unit": "rows/s", "gpu": n_rows / (ms * 1e-3), "gpu_ms": ms,
                    "hbm_write_GBs": n_rows * (lay.n + (8 if kind == 2 else 0)) / (ms * 1e-3) / 1e9,
                    "cpu": {"value": n_cpu / dt, "cores": cores, "kind": "port", "sample": f"{n_cpu} rows"},
                    "check": f"first {n_cpu} rows identical to the oracle: {same}"})
    # ---- f4 statistics: count the ancestral table where it lies (sample -> count never leaves HBM)
    _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(no_ev, _lib.i8p), C.c_int32(0), C.c_int32(1), C.c_uint64(0),
                               C.c_uint64(n_rows), C.c_uint64(3), C.c_void_p(st.data_ptr()), None, None))
    cnt = torch.empty(int(lay.cpt_off[-1]), dtype=torch.int64, device="cuda:0")

    def fn_stats():
        _lib.check(L.bn_statistics_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_rows), C.c_uint64(n_rows),
                                       C.c_void_p(cnt.data_ptr())))
    ms = ev_time(torch, fn_stats)
    n_cpu = min(n_rows, 1 << 21)
    host_tab = np.ascontiguousarray(st[:, :n_cpu].cpu().numpy())
    capi.statistics(on, host_tab)
    t0 = time.perf_counter()
    cref = capi.statistics(on, host_tab)
    dt = time.perf_counter() - t0
    _lib.check(L.bn_statistics_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_cpu), C.c_uint64(n_rows),
                                   C.c_void_p(cnt.data_ptr())))
    same = bool(np.array_equal(cnt.cpu().numpy(), cref))
    out.append({"row": "f4", "api": "statistics(bn, data): all nodes' count tables in one pass",
                "cite": "src/DiscreteBayesNet/discrete_bayes_net.jl:151-172,221-231",
                "config": f"C2 net, device-resident {n_rows} rows x {lay.n} nodes (uint8) from rand(bn, n)", "unit": "rows/s",
                "gpu": n_rows / (ms * 1e-3), "gpu_ms": ms, "hbm_read_GBs": n_rows * lay.n / (ms * 1e-3) / 1e9,
                "cpu": {"value": n_cpu / dt, "cores": cores, "kind": "port", "sample": f"{n_cpu} rows (OpenMP over nodes)"},
                "check": f"counts of the first {n_cpu} rows identical to the oracle: {same}"})
    # ---- f5 logpdf(bn, df): the table's log-likelihood = its counts . log CPTs (counting kernel + one small reduction)
    ll = torch.empty(1, dtype=torch.float64, device="cuda:0")

    def fn_logpdf():
        _lib.check(L.bn_logpdf_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_rows), C.c_uint64(n_rows),
                                   C.c_void_p(ll.data_ptr())))
    ms = ev_time(torch, fn_logpdf)
    capi.logpdf_table(on, host_tab)
    t0 = time.perf_counter()
    lref = capi.logpdf_table(on, host_tab)
    dt = time.perf_counter() - t0
    _lib.check(L.bn_logpdf_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_cpu), C.c_uint64(n_rows),
                               C.c_void_p(ll.data_ptr())))
    lgot = float(ll.cpu()[0])
    out.append({"row": "f5", "api": "logpdf(bn, df): data log-likelihood of a table",
                "cite": "src/ProbabilisticGraphicalModels/ProbabilisticGraphicalModels.jl:83-104, src/bayes_nets.jl:322-328",
                "config": f"C2 net, device-resident {n_rows} rows x {lay.n} nodes (uint8) from rand(bn, n)", "unit": "rows/s",
                "gpu": n_rows / (ms * 1e-3), "gpu_ms": ms, "hbm_read_GBs": n_rows * lay.n / (ms * 1e-3) / 1e9,
                "cpu": {"value": n_cpu / dt, "cores": 1, "kind": "port",
                        "sample": f"{n_cpu} rows (row-by-row sum like the reference, one thread)"},
                "check": f"first {n_cpu} rows: device {lgot!r} vs oracle {lref!r}, rel diff {abs(lgot - lref) / abs(lref):.2e}"})
    del st, w, ok, cnt, ll

    # ---- a8 gibbs_sample: 4096 chains x 16 kept samples, burn-in 50 sweeps, thinning 1 (C2 net)
    nch, ns, bi, th = (4096, 16, 50, 1) if not a.quick else (256, 4, 10, 1)
    buf = np.zeros((lay.n, nch * ns), np.uint8)
    t0 = time.perf_counter()
    _lib.check(L.bn_gibbs_sample(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), C.c_uint64(0), C.c_uint64(nch), C.c_int64(ns),
                                 C.c_int64(bi), C.c_int64(th), None, C.c_uint64(5), None, _lib.ptr(buf, _lib.u8p)))
    t0 = time.perf_counter()
    _lib.check(L.bn_gibbs_sample(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), C.c_uint64(0), C.c_uint64(nch), C.c_int64(ns),
                                 C.c_int64(bi), C.c_int64(th), None, C.c_uint64(5), None, _lib.ptr(buf, _lib.u8p)))
    dt_gpu = time.perf_counter() - t0
    n_free = int((ev < 0).sum())
    sweeps = bi + ns * (th + 1)
    cch = 64
    t0 = time.perf_counter()
    cbuf = capi.gibbs_sample(on, ev, cch, ns, bi, thinning=th, seed=5)
    dt_cpu = time.perf_counter() - t0
    same = bool(np.array_equal(buf[:, :cch * ns], cbuf))
    out.append({"row": "a8", "api": "gibbs_sample(bn, nsamples, burn_in; thinning)", "cite": "src/gibbs.jl:183-245,289-374",
                "config": f"C2 net, {nch} chains x {ns} samples, burn-in {bi} sweeps, thinning {th}, random sweep order, "
                          "host call incl. order upload + table download", "unit": "node-updates/s",
                "gpu": nch * sweeps * n_free / dt_gpu, "gpu_ms": dt_gpu * 1e3,
                "cpu": {"value": cch * sweeps * n_free / dt_cpu, "cores": cores, "kind": "port", "sample": f"{cch} chains"},
                "check": f"first {cch} chains identical to the oracle: {same}"})

    # ---- a7 Gibbs infer (C4) -- tools/gibbs_bench.py
    import gibbs_bench
    g = gibbs_bench.run(bn, torch, n_chains=65536 if not a.quick else 4096)
    out.append({"row": "a7", "api": "infer(GibbsSamplingNodewise(2000, 1000, 5), bn, query; evidence), 65536 chains",
                "cite": "src/Inference/gibbs.jl:66-145", "config": "C4: 200-node net, 20 evidence", "unit": "node-updates/s",
                "gpu": g["node_updates_per_s"], "gpu_ms": g["ms"], "cpu": g["cpu_baseline"],
                "check": f"max |posterior - LW(2e9)| = {g['max_abs_err']:.4f} (see DESIGN.md 3.3); counts == oracle in tests"})

    # ---- a10 / a11 / a12: factor ops vs the C oracle on the same 2^22-entry operands (GB/s, algorithmic bytes)
    k = 22 if not a.quick else 18
    rng = np.random.default_rng(0)
    names = [f"pb{i}" for i in range(k)]
    A = rng.random((2,) * k)
    B = rng.random((2,) * (k // 2))
    fa, fb = bn.Factor(names, A), bn.Factor(names[0::2], B)
    N = 1 << k
    for row, name, cite, gfn, cfn, nbytes in [
            ("a10", "phi1 * phi2 (2^%d x 2^%d interleaved)" % (k, k // 2), "src/Factors/factors_dims.jl:185-234,269",
             lambda: fa * fb, lambda: capi.factor_product(list(range(k)), A, list(range(0, k, 2)), B), 8 * (2 * N + (1 << (k // 2)))),
            ("a11", "sum(phi, middle dim) of 2^%d" % k, "src/Factors/factors_dims.jl:60-85,100",
             lambda: fa.sum(names[k // 2]), lambda: capi.factor_sum(list(range(k)), A, [k // 2]), 8 * (N + N // 2)),
            ("a12", "normalize!(phi) of 2^%d" % k, "src/Factors/factors_dims.jl:45-57",
             lambda: fa.copy().normalize_(), lambda: capi.factor_normalize(A.copy()), 8 * 3 * N)]:
        ms = ev_time(torch, gfn)
        t0 = time.perf_counter()
        cres = cfn()
        dt = time.perf_counter() - t0
        gres = gfn().potential
        cval = cres[1] if isinstance(cres, tuple) else cres
        same = bool(np.array_equal(gres, np.asarray(cval).reshape(gres.shape, order="F"))) if row != "a12" else \
            bool(np.allclose(gres, np.asarray(cval).reshape(gres.shape, order="F"), rtol=1e-12, atol=0))
        out.append({"row": row, "api": name, "cite": cite, "config": "single Python-level op incl. result allocation "
                    "(kernel-rate figures at 2^24: bench.py factor_roofline)", "unit": "GB/s",
                    "gpu": nbytes / (ms * 1e-3) / 1e9, "gpu_ms": ms,
                    "cpu": {"value": nbytes / dt / 1e9, "cores": cores, "kind": "port", "sample": "same operands"},
                    "check": ("bit-identical to the oracle: " if row != "a12" else "within 1e-12 of the oracle: ") + str(same)})

    # ---- a9 data movement: permutedims (rotate the dims by 7: a strided gather) and push! (repeat along a new dim)
    from oracle.factors_np import NpFactor
    rot = list(range(8, k + 1)) + list(range(1, 8))
    nf_a = NpFactor(names, np.asfortranarray(A))
    for row, name, cite, gfn, cfn, nbytes in [
            ("a9p", "permutedims(phi, rotate by 7) of 2^%d" % k, "src/Factors/factors_main.jl:160-166",
             lambda: fa.permutedims(rot), lambda: nf_a.permutedims([r - 1 for r in rot]).potential, 8 * 2 * N),
            ("a9d", "push!(phi, dim, 4) of 2^%d" % k, "src/Factors/factors_main.jl:148-158",
             lambda: fa.copy().push_("pbq", 4), lambda: nf_a.push("pbq", 4).potential, 8 * (N + 4 * N))]:
        ms = ev_time(torch, gfn)
        if row == "a9d":
            ms -= ev_time(torch, lambda: fa.copy())  # the copy that keeps fa intact is not part of push!
        t0 = time.perf_counter()
        cval = cfn()
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(gfn().potential, cval))
        out.append({"row": row, "api": name, "cite": cite, "config": "single Python-level op incl. result allocation",
                    "unit": "GB/s", "gpu": nbytes / (ms * 1e-3) / 1e9, "gpu_ms": ms,
                    "cpu": {"value": nbytes / dt / 1e9, "cores": 1, "kind": "port", "sample": "same operand (numpy oracle)"},
                    "check": "identical to the oracle: " + str(same)})

    # ---- a13 ExactInference end to end on the C5 net (one 2^24-entry CPT)
    kk = 23 if not a.quick else 15
    lay5 = c5_layout(bn, k=kk)
    dn5 = DeviceNet(lay5, 0)
    ev5 = np.full(lay5.n, -1, np.int8)
    ev5[3] = 1
    q5 = np.array([lay5.n - 1], np.int32)
    res = {}
    for fused in (1, 0):
        o = np.zeros(2)
        od, oc = np.zeros(1, np.int32), np.zeros(1, np.int32)

        def fn():
            _lib.check(L.bn_exact_infer(_lib.h64(dn5.handle), _lib.ptr(ev5, _lib.i8p), _lib.ptr(q5, _lib.i32p), C.c_int32(1),
                                        C.c_int32(fused), _lib.ptr(od, _lib.i32p), _lib.ptr(oc, _lib.i32p), _lib.ptr(o, _lib.f64p)))
        fn(); fn()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        res[fused] = (float(np.median(ts)), o.copy())
    t0 = time.perf_counter()
    ref = exact_infer(onet(lay5), [lay5.names[-1]], {lay5.names[3]: 1})
    dt = time.perf_counter() - t0
    refp = ref.permuted_to([lay5.names[-1]])
    err = float(np.max(np.abs(res[1][1] - refp) / refp))
    out.append({"row": "a13", "api": "infer(ExactInference(), bn, query; evidence)", "cite": "src/Inference/exact.jl:6-33",
                "config": f"C5: 30 binary nodes, one with {kk} parents (CPT 2^{kk + 1} doubles), 1 evidence, 1 query; host call "
                          "incl. ordering, all kernels, result download", "unit": "inferences/s",
                "gpu": 1.0 / res[1][0], "gpu_ms": res[1][0] * 1e3, "gpu_unfused_ms": res[0][0] * 1e3,
                "cpu": {"value": 1.0 / dt, "cores": 1, "kind": "port", "sample": "oracle numpy VE (pairwise join + sum), same net"},
                "check": f"max rel err vs oracle VE = {err:.2e} (bar 1e-12); fused == unfused: "
                         f"{bool(np.allclose(res[1][1], res[0][1], rtol=1e-12, atol=0))}"})

    # ---- f3 LoopyBelief batched over evidence sets (C2 net; tools/loopy_bench.py holds the case)
    import loopy_bench
    n_inst, n_it = (16384, 50) if not a.quick else (1184, 10)
    lb = loopy_bench.run(bn, torch, n_inst, n_it, n_check=2)
    out.append({"row": "f3", "api": "infer(LoopyBelief(), bn, query; evidence), batched over (query, evidence) pairs",
                "cite": "src/Inference/loopy_belief.jl:24-218",
                "config": f"C2 net, {n_inst} (query, 10 evidence nodes) pairs x {n_it} iterations (tol < 0), one launch, "
                          "device-resident evidence rows", "unit": "instance-iterations/s",
                "gpu": lb["instance_iterations_per_s"], "gpu_ms": lb["ms"], "message_updates_per_s": lb["message_updates_per_s"],
                "cpu": lb["cpu_baseline"], "check": f"beliefs identical to the numpy oracle (2 pairs): {lb['identical_to_oracle']}, to the C oracle "
                         f"({lb['identical_to_c_oracle']['pairs']} pairs): {lb['identical_to_c_oracle']['identical']}"})

    for r in out:
        cpu = r["cpu"]["value"]
        print(f"{r['row']:4s} {r['api'][:62]:62s} gpu {r['gpu']:.4g} {r['unit']}  cpu {cpu:.4g} ({r['cpu']['cores']} cores)  x{r['gpu'] / cpu:.0f}  | {r['check']}")
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    return out


if __name__ == "__main__":
    main()
