#!/usr/bin/env python
"""statistics(bn, data) on a device-resident table: rows/s and fraction of the HBM roofline (n_nodes bytes per row),
for the row-lane kernel and (BN_STATS_NODELANES=1) the node-lane fallback, with a parity check against the oracle.

    python tools/stats_bench.py [--rows 16777216] [--json out.json]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 24)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import torch
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import device_net
    from oracle import capi
    import bench
    L = _lib.lib()
    net, lay, query, evidence = bench.build_case(bn)
    dn = device_net(net, 0)
    n_rows = a.rows
    bn.set_lw_mode("specialized", True)
    st = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
    cnt = torch.empty(int(lay.cpt_off[-1]), dtype=torch.int64, device="cuda:0")
    no_ev = np.full(lay.n, -1, np.int8)
    _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(no_ev, _lib.i8p), C.c_int32(0), C.c_int32(1), C.c_uint64(0),
                               C.c_uint64(n_rows), C.c_uint64(3), C.c_void_p(st.data_ptr()), None, None))

    def fn():
        _lib.check(L.bn_statistics_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_rows), C.c_uint64(n_rows),
                                       C.c_void_p(cnt.data_ptr())))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = float(np.median(ms))
    n_cpu = min(n_rows, 1 << 20)
    on = bench.oracle_net(lay)
    _lib.check(L.bn_statistics_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_cpu), C.c_uint64(n_rows),
                                   C.c_void_p(cnt.data_ptr())))
    same = bool(np.array_equal(cnt.cpu().numpy(), capi.statistics(on, np.ascontiguousarray(st[:, :n_cpu].cpu().numpy()))))
    hbm, src = bench.peaks()
    gbs = n_rows * lay.n / (ms * 1e-3) / 1e9
    res = {"kernel": "node-lane (fallback)" if os.environ.get("BN_STATS_NODELANES") else "row-lane", "rows": n_rows, "n_nodes": int(lay.n),
           "bins": int(lay.cpt_off[-1]), "ms": ms, "rows_per_s": n_rows / (ms * 1e-3), "hbm_read_GBs": gbs, "frac_of_hbm": gbs / hbm,
           "peak": hbm, "peak_source": src, "counts_identical_to_oracle": same, "warps_env": os.environ.get("BN_STATS_WARPS")}
    print(json.dumps(res))
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
