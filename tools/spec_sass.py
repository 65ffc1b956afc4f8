#!/usr/bin/env python
"""SASS evidence for the NVRTC (network-specialised) kernels, made on the host: generate the source for the bench
cases, compile it for sm_100a through the library's own cubin cache (bn_lw_spec_source / bn_gibbs_spec_source, no GPU
needed), run `cuobjdump -sass` on the cached cubin and print a mnemonic histogram + resource usage.

    python tools/spec_sass.py > profiles/r2_spec_sass.txt
"""
import collections
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def extract_cubin(cache_file):
    """A cache entry is [magic u64 | source bytes u64 | cubin bytes u64 | source | cubin] (csrc/specialize.cu)."""
    raw = open(cache_file, "rb").read()
    magic, n_src, n_cub = np.frombuffer(raw[:24], np.uint64)
    assert int(magic) == 0x314e4942554342 and len(raw) == 24 + int(n_src) + int(n_cub)
    out = cache_file + ".elf"
    open(out, "wb").write(raw[24 + int(n_src):])
    return out


def histogram(cubin_path):
    cubin_path = extract_cubin(cubin_path)
    sass = subprocess.run(["cuobjdump", "-sass", cubin_path], capture_output=True, text=True, check=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops[m.group(1)] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", cubin_path], capture_output=True, text=True).stdout
    return ops, res


def report(title, ops, res, want):
    total = sum(ops.values())
    base = collections.Counter()
    for k, v in ops.items():
        base[k.split(".")[0]] += v
    print(f"== {title}: {total} SASS instructions")
    for line in res.splitlines():
        if "REG" in line or "Function" in line:
            print("   " + line.strip())
    print("   top opcodes: " + ", ".join(f"{k} {v}" for k, v in base.most_common(14)))
    for w in want:
        hits = {k: v for k, v in ops.items() if k.startswith(w)}
        print(f"   {w:8s}: {sum(hits.values())}  {dict(sorted(hits.items())) if hits else ''}")
    print()


def main():
    cache = tempfile.mkdtemp(prefix="bn_sass_")
    os.environ["BN_CUBIN_CACHE"] = cache
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    import bench
    import gibbs_bench
    from dump_lw_spec import spec_source
    print("# SASS of the NVRTC kernels generated for the bench cases (tools/spec_sass.py; host-only, sm_100a cubins\n"
          "# from the library's cubin cache, cuobjdump -sass).  UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier,\n"
          "# IMAD.WIDE = the 32x32->64 multiplies of Philox4x32-10, LDS = table-row loads, LEA.HI / SHF = threshold compares.\n")

    def newest():
        fs = [os.path.join(cache, f) for f in os.listdir(cache) if f.endswith(".cubin")]
        return max(fs, key=os.path.getmtime)

    net, lay, query, evidence = bench.build_case(bn)
    for prune in (False, True):
        src, cub, live = spec_source(lay, evidence, query, prune=prune)
        ops, res = histogram(newest())
        report(f"lw_spec, C2 net (100 nodes, 10 evidence, query {query}), {'barren nodes pruned' if prune else 'all nodes sampled (bench headline)'}: "
               f"{live} live nodes, cubin {cub} bytes", ops, res, ["UBLKCP", "SYNCS", "IMAD.WIDE", "LDS", "LEA.HI", "SHF", "ATOM", "RED", "DADD", "DMUL"])
    # Gibbs C4
    gnet = bn.rand_discrete_bn(200, 3, 4, seed=0)
    glay = bn.flatten(gnet)
    rng = np.random.default_rng(4000)
    picks = [str(x) for x in rng.choice(glay.names, size=21, replace=False)]
    gq, gev = picks[:1], {e: 1 for e in picks[1:]}
    ev, q = glay.evidence_vector(gev), glay.query_indices(gq)
    for threads in (256, 64):
        n_len, cub = C.c_int64(0), C.c_int64(0)
        args = [C.c_int32(glay.n), _lib.ptr(glay.card, _lib.i32p), _lib.ptr(glay.par_ptr, _lib.i32p),
                _lib.ptr(glay.par_idx, _lib.i32p), _lib.ptr(glay.cpt_off, _lib.i64p), _lib.ptr(glay.cpt, _lib.f64p),
                _lib.ptr(ev, _lib.i8p), _lib.ptr(q, _lib.i32p), C.c_int32(q.size), C.c_int32(0), C.c_int32(threads)]
        _lib.check(_lib.lib().bn_gibbs_spec_source(*args, None, C.c_int64(0), C.byref(n_len), C.byref(cub)))
        ops, res = histogram(newest())
        report(f"gibbs_spec, C4 net (200 nodes, 20 evidence), Nodewise, {threads}-thread CTAs: cubin {cub.value} bytes, source "
               f"{n_len.value} chars", ops, res, ["LDS", "STS", "LDG", "DMUL", "DADD", "DSETP", "BAR", "CALL", "IMAD.WIDE"])


if __name__ == "__main__":
    main()
