#!/usr/bin/env python
"""rand(bn, n) -> statistics on one GPU, event-timed as a pair and separately (C2 net, 2^24 rows per step)."""
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
    import bench
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.distributed import ShardedTableStatistics
    net, lay, query, evidence = bench.build_case(bn)
    bn.set_lw_mode("specialized", False)
    _lib.set_stream(torch.cuda.current_stream(0).cuda_stream)
    n = 1 << 24
    job = ShardedTableStatistics(net, n, seed=1, device=0)
    for _ in range(3):
        job.step()
    torch.cuda.synchronize()

    def timed(fn, reps=5):
        ms = []
        for r in range(reps):
            job.seed = 10 + r
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))
    pair, samp, cnt = timed(job.step), timed(job.sample), timed(job.launch)
    import time

    def host(fn, reps=20):
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
        return float(np.median(ts)) * 1e3
    print(json.dumps({"host_ms_sample_call": round(host(job.sample), 4), "host_ms_count_call": round(host(job.launch), 4)}))
    print(json.dumps({"rows": n, "pair_ms": round(pair, 4), "pair_rows_per_s": n / (pair * 1e-3), "sample_ms": round(samp, 4),
                      "sample_rows_per_s": n / (samp * 1e-3), "count_ms": round(cnt, 4), "count_rows_per_s": n / (cnt * 1e-3)}))
