#!/usr/bin/env python
"""Host-side cost of one factor op through the C ABI (tiny operands, so the device time is negligible and the stream
never backs up): what the caller pays per `*` / `sum` besides the kernel."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    import torch
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    L = _lib.lib()
    a = bn.Factor(["h0", "h1", "h2"], np.random.rand(2, 2, 2))
    b = bn.Factor(["h1", "h3"], np.random.rand(2, 2))
    ids = np.array([bn.factors.var_id("h1")], np.int32)
    n = 20000

    def loop(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6

    def f_sum():
        out = _lib.h64(0)
        L.bn_factor_sum(_lib.h64(a.handle), _lib.ptr(ids, _lib.i32p), C.c_int32(1), C.byref(out))
        L.bn_factor_free(out)

    def f_prod():
        out = _lib.h64(0)
        L.bn_factor_product(_lib.h64(a.handle), _lib.h64(b.handle), C.byref(out))
        L.bn_factor_free(out)

    def f_noop():
        nd = C.c_int32(0)
        L.bn_factor_ndims(_lib.h64(a.handle), C.byref(nd))

    for name, fn in [("ctypes call floor (bn_factor_ndims)", f_noop), ("sum + free", f_sum), ("product + free", f_prod)]:
        fn()
        h, tot = loop(fn)
        print(f"{name:40s} host {h:6.2f} us/op   incl. drain {tot:6.2f} us/op")
    # short bursts (the launch queue never fills, so this is the host's own cost per op)
    for name, fn in [("sum + free, bursts of 200", f_sum), ("product + free, bursts of 200", f_prod)]:
        ts = []
        for _ in range(30):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
            ts.append((time.perf_counter() - t0) / 200 * 1e6)
        print(f"{name:40s} host {np.median(ts):6.2f} us/op")
