#!/usr/bin/env python
"""bn_exact_infer on the C5 net (tools/path_bench.py c5_layout): wall time per call, kernels launched per call, and
(under ncu --metrics gpu__time_duration.sum) the per-kernel device time."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    import torch
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.device_layout import DeviceNet
    from path_bench import c5_layout
    kk = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    lay5 = c5_layout(bn, k=kk)
    dn5 = DeviceNet(lay5, 0)
    L = _lib.lib()
    ev5 = np.full(lay5.n, -1, np.int8)
    ev5[3] = 1
    q5 = np.array([lay5.n - 1], np.int32)
    o, od, oc = np.zeros(2), np.zeros(1, np.int32), np.zeros(1, np.int32)

    def fn(fused=1):
        _lib.check(L.bn_exact_infer(_lib.h64(dn5.handle), _lib.ptr(ev5, _lib.i8p), _lib.ptr(q5, _lib.i32p), C.c_int32(1),
                                    C.c_int32(fused), _lib.ptr(od, _lib.i32p), _lib.ptr(oc, _lib.i32p), _lib.ptr(o, _lib.f64p)))
    fn(); fn()
    l0 = bn.launch_count()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print("fused: ms/call", round(float(np.median(ts)) * 1e3, 4), "launches/call", (bn.launch_count() - l0) / 10, o)
