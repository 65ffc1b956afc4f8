#!/usr/bin/env python
"""Factor `*` / sum / normalize! / fused-eliminate sweep (BASELINE C5 shapes) with two timings per case:
  call  : one op per CUDA-event pair, L2 flushed before it (includes the host's launch latency)
  batch : `reps` ops back to back inside one event pair, rotating over `nrot` distinct operand sets so that
          consecutive ops never touch the same bytes (working set >> 126 MB L2); this is the kernel-rate figure.
    python tools/factor_bench.py [--k 24] [--reps 12] [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def build_cases(bn, k, nrot, rng):
    names = [f"w{i}" for i in range(k + 2)]

    def mk(dims):
        return bn.Factor([names[i] for i in dims], rng.random(1 << len(dims)).reshape((2,) * len(dims), order="F"))

    N = 1 << k
    perm = [int(x) for x in rng.permutation(k)]
    sets = []
    for _ in range(nrot):
        sets.append({
            "A": mk(range(k)), "B12": mk(range(0, k, 2)), "A23": mk(range(k - 1)), "B23": mk(range(1, k)),
            "Bp": mk(perm), "Eb": mk([0] + list(range(k - 12, k - 1))), "Ec": mk([0] + list(range(1, 12))),
        })
    # ops go straight through the C ABI on handles (what a `ccall` does); wrapping every result in a Python Factor
    # costs ~20 us of interpreter time per op, which at 35-60 us per kernel would make the HOST the bottleneck
    from bayesnets_jl_b200 import _lib
    import ctypes as C
    L = _lib.lib()

    class Raw:
        """owns one bn_factor_t; freed (stream-ordered) when dropped"""
        __slots__ = ("h",)

        def __init__(self, h):
            self.h = h

        def __del__(self):
            L.bn_factor_free(self.h)

    def ids(*dd):
        a = np.array([bn.factors.var_id(names[i]) for i in dd], np.int32)
        return a, _lib.ptr(a, _lib.i32p), C.c_int32(len(dd))

    def product(a, b):
        out = _lib.h64(0)
        _lib.check(L.bn_factor_product(_lib.h64(a.handle), _lib.h64(b.handle), C.byref(out)))
        return Raw(out)

    def fsum(a, idp):
        out = _lib.h64(0)
        _lib.check(L.bn_factor_sum(_lib.h64(a.handle), idp[1], idp[2], C.byref(out)))
        return Raw(out)

    def elim(fs, idp):
        hs = (_lib.h64 * len(fs))(*[f.handle for f in fs])
        out = _lib.h64(0)
        _lib.check(L.bn_factor_eliminate(hs, C.c_int32(len(fs)), idp[1], idp[2], C.byref(out)))
        return Raw(out)

    def norm(a):
        _lib.check(L.bn_factor_normalize(_lib.h64(a.handle), C.c_int32(1)))
        return None

    def permute(a, perm):
        out = _lib.h64(0)
        _lib.check(L.bn_factor_permutedims(_lib.h64(a.handle), _lib.ptr(perm, _lib.i32p), C.c_int32(len(perm)), C.byref(out)))
        return Raw(out)

    def push(a, dim, length):
        out = _lib.h64(0)
        _lib.check(L.bn_factor_push(_lib.h64(a.handle), C.c_int32(dim), C.c_int32(length), C.byref(out)))
        return Raw(out)

    rot = np.array(list(range(7, k)) + list(range(7)), np.int32)  # out dim i = in dim (i + 7) mod k
    qdim = bn.factors.var_id(names[k + 1])
    d0, dk, dm, d4 = ids(0), ids(k - 1), ids(k // 2), ids(1, 5, 11, k - 2)
    cases = [
        ("P1 2^%d x 2^%d interleaved" % (k, k // 2), lambda s: product(s["A"], s["B12"]), 8 * (2 * N + (1 << (k // 2)))),
        ("P2 dims1..k-1 x dims2..k", lambda s: product(s["A23"], s["B23"]), 8 * (N + N)),
        ("P3 2^%d x 2^%d permuted" % (k, k), lambda s: product(s["A"], s["Bp"]), 8 * 3 * N),
        ("S1 sum innermost", lambda s: fsum(s["A"], d0), 8 * (N + N // 2)),
        ("S2 sum outermost", lambda s: fsum(s["A"], dk), 8 * (N + N // 2)),
        ("S3 sum middle", lambda s: fsum(s["A"], dm), 8 * (N + N // 2)),
        ("S4 sum 4 dims", lambda s: fsum(s["A"], d4), 8 * (N + N // 16)),
        ("N1 normalize", lambda s: norm(s["A"]), 8 * 3 * N),
        ("T1 permutedims rotate by 7", lambda s: permute(s["A"], rot), 8 * 2 * N),
        ("D1 push! 2^%d x 2 -> 2^%d" % (k - 1, k), lambda s: push(s["A23"], qdim, 2), 8 * (N // 2 + N)),
        ("E1 fused eliminate 3 factors", lambda s: elim([s["A"], s["Eb"], s["Ec"]], d0),
         8 * (N + 2 * (1 << 12) + N // 2)),
    ]
    return sets, cases


def run(bn, torch, device=0, k=24, reps=24, nrot=3, only=None):
    from bayesnets_jl_b200 import _lib
    rng = np.random.default_rng(0)
    sets, cases = build_cases(bn, k, nrot, rng)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=f"cuda:{device}")
    _lib.set_stream(torch.cuda.current_stream(device).cuda_stream)
    out = []
    for name, fn, nbytes in cases:
        if only and not name.startswith(only):
            continue
        keep = []  # warm-up with the batch loop's own allocation pattern (pool growth, code load)
        for it in range(2 * nrot + 2):
            keep.append(fn(sets[it % nrot]))
            if len(keep) > 2:
                keep.pop(0)
        del keep
        torch.cuda.synchronize()
        call = []
        for it in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(sets[it % nrot])
            e1.record()
            torch.cuda.synchronize()
            del r
            call.append(e0.elapsed_time(e1))
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        keep = []
        import time
        t0 = time.perf_counter()
        e0.record()
        for it in range(reps):
            keep.append(fn(sets[it % nrot]))
            if len(keep) > 2:
                keep.pop(0)  # frees go back to the stream-ordered pool; no sync
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / reps
        torch.cuda.synchronize()
        del keep
        batch_ms = e0.elapsed_time(e1) / reps
        out.append({"case": name, "bytes": nbytes, "call_ms": round(float(np.mean(call)), 4),
                    "call_GBs": round(nbytes / np.mean(call) / 1e6, 1), "batch_ms": round(batch_ms, 4),
                    "batch_GBs": round(nbytes / batch_ms / 1e6, 1), "host_ms_per_op": round(host_ms, 4)})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=24)
    ap.add_argument("--reps", type=int, default=24)
    ap.add_argument("--only", default=None)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import torch
    import bayesnets_jl_b200 as bn
    res = run(bn, torch, k=a.k, reps=a.reps, only=a.only)
    peak = 6546.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    for r in res:
        r["batch_frac"] = round(r["batch_GBs"] / peak, 3)
        print(f"{r['case']:34s} call {r['call_ms']:.4f} ms {r['call_GBs']:8.1f} GB/s | batch {r['batch_ms']:.4f} ms "
              f"{r['batch_GBs']:8.1f} GB/s  ({r['batch_frac']:.3f} of {peak:.0f})  host {r['host_ms_per_op']:.4f} ms/op")
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)
