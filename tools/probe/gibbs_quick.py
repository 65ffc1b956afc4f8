#!/usr/bin/env python
"""Time the specialised Gibbs inference kernel on the C4 case (65 536 chains) and at 1 M chains; prints ms, node
updates/s and the raw counts (identical across kernel variants).  python tools/probe/gibbs_quick.py [chains ...]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

if __name__ == "__main__":
    import torch
    import bayesnets_jl_b200 as bn
    import gibbs_bench
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import device_net
    nn = int(os.environ.get("GQ_NODES", "200"))
    net, lay, query, evidence = gibbs_bench.build_case(bn, n_nodes=nn, n_ev=nn // 10)
    bn.set_gibbs_mode(os.environ.get("BN_GIBBS_KERNEL", "specialized"))
    dn = device_net(net, 0)
    ev, q = lay.evidence_vector(evidence), lay.query_indices(query)
    ncell = int(np.prod([lay.card[i] for i in q]))
    out = torch.zeros(ncell, dtype=torch.float64, device="cuda:0")
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    L = _lib.lib()
    for n_chains in [int(x) for x in (sys.argv[1:] or ["65536", "1048576"])]:
        def launch(seed):
            out.zero_()
            _lib.check(L.bn_gibbs_infer_dev(_lib.h64(dn.handle), _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p),
                                            C.c_int32(q.size), C.c_uint64(0), C.c_uint64(n_chains), C.c_int64(2000),
                                            C.c_int64(1000), C.c_int64(5), C.c_int32(0), C.c_uint64(seed),
                                            C.c_void_p(out.data_ptr())))
        for w in range(3):
            launch(100 + w)
        torch.cuda.synchronize()
        import time
        ms, host = [], []
        for r in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            launch(r)
            torch.cuda.synchronize()
            e0.record()
            t0 = time.perf_counter()
            launch(r)
            host.append((time.perf_counter() - t0) * 1e3)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = float(np.median(ms))
        print(json.dumps({"nodes": nn, "chains": n_chains, "ms": round(m, 4), "host_ms": round(float(np.median(host)), 4), "cycles_per_update_per_warp": round(m * 1e-3 * 1.965e9 / 2000, 1), "updates_per_s": n_chains * 2000 / (m * 1e-3),
                          "counts_seed4": [int(x) for x in out.cpu().numpy()]}))
