#!/usr/bin/env python
"""One bn_statistics_dev launch on a device-resident C2 table inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --set full --import-source on -o gpurun_out/prof_stats python tools/stats_profile.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    import torch
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import device_net
    import bench
    n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
    net, lay, query, evidence = bench.build_case(bn)
    dn = device_net(net, 0)
    L = _lib.lib()
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    st = torch.empty((lay.n, n_rows), dtype=torch.uint8, device="cuda:0")
    cnt = torch.empty(int(lay.cpt_off[-1]), dtype=torch.int64, device="cuda:0")
    no_ev = np.full(lay.n, -1, np.int8)
    _lib.check(L.bn_sample_dev(_lib.h64(dn.handle), _lib.ptr(no_ev, _lib.i8p), C.c_int32(0), C.c_int32(1), C.c_uint64(0),
                               C.c_uint64(n_rows), C.c_uint64(3), C.c_void_p(st.data_ptr()), None, None))

    def run():
        _lib.check(L.bn_statistics_dev(_lib.h64(dn.handle), C.c_void_p(st.data_ptr()), C.c_uint64(n_rows), C.c_uint64(n_rows),
                                       C.c_void_p(cnt.data_ptr())))
    run(); run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print("rows/s", n_rows / (e0.elapsed_time(e1) * 1e-3))
