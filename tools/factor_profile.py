#!/usr/bin/env python
"""Run each factor sweep case exactly once inside a cudaProfilerStart/Stop range (after warm-up), for
    ncu --profile-from-start off --set full -o gpurun_out/prof_factor python tools/factor_profile.py"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    import torch
    import bayesnets_jl_b200 as bn
    from factor_bench import build_cases
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    sets, cases = build_cases(bn, k, 1, np.random.default_rng(0))
    for name, fn, nbytes in cases:
        for _ in range(2):
            r = fn(sets[0]); del r
    torch.cuda.synchronize()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for name, fn, nbytes in cases:
        r = fn(sets[0]); del r
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", [c[0] for c in cases])
