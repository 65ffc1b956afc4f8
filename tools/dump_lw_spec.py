#!/usr/bin/env python
"""Emit (and NVRTC-compile, no GPU needed) the network-specialised LW kernel source for bench.py's C2 case.

    python tools/dump_lw_spec.py [--all-nodes] [--out /tmp/lw_spec.cu]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bayesnets_jl_b200 as bn  # noqa: E402
from bayesnets_jl_b200 import _lib  # noqa: E402
import bench  # noqa: E402


def spec_source(lay, evidence, query, prune=True, compile=True):
    ev, q = lay.evidence_vector(evidence), lay.query_indices(query)
    n_len, cub, nl = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    par_idx = lay.par_idx if lay.par_idx.size else np.zeros(1, np.int32)
    args = [C.c_int32(lay.n), _lib.ptr(lay.card, _lib.i32p), _lib.ptr(lay.par_ptr, _lib.i32p),
            _lib.ptr(par_idx, _lib.i32p), _lib.ptr(lay.cpt_off, _lib.i64p), _lib.ptr(lay.cpt, _lib.f64p),
            _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p), C.c_int32(q.size), C.c_int32(1 if prune else 0)]
    _lib.check(_lib.lib().bn_lw_spec_source(*args, None, C.c_int64(0), C.byref(n_len), None, None))
    buf = C.create_string_buffer(n_len.value + 1)
    _lib.check(_lib.lib().bn_lw_spec_source(*args, buf, C.c_int64(n_len.value + 1), C.byref(n_len),
                                            C.byref(cub) if compile else None, C.byref(nl)))
    return buf.value.decode(), cub.value, nl.value


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--all-nodes", action="store_true")
    ap.add_argument("--out", default="/tmp/lw_spec.cu")
    a = ap.parse_args()
    net, lay, query, evidence = bench.build_case(bn)
    src, cub, nl = spec_source(lay, evidence, query, prune=not a.all_nodes)
    open(a.out, "w").write(src)
    print(f"{a.out}: {len(src)} chars, {nl} live nodes of {lay.n}, cubin {cub} bytes")
