#!/usr/bin/env python
"""Multi-GPU check + timing of the path inside libbn_cuda.so (csrc/comm.cu), no torch anywhere.

    python tools/multi_gpu_check.py --all 0,1,2,3            one process drives the listed GPUs (ncclCommInitAll)
    torchrun --nproc-per-node N ... tools/multi_gpu_check.py  one process per GPU (ncclCommInitRank, id via a file)

LW on the C2 net and Gibbs C4 through the comm-bearing net, compared with the same job on ONE GPU: LW counts to 1e-12,
Gibbs counts identical; row-sharded table statistics identical.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", default=None, help="comma-separated device list: single-process mode")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--particles", type=float, default=None, help="LW particles per GPU")
    a = ap.parse_args()
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200.distributed import (Communicator, ShardedGibbs, ShardedLikelihoodWeighting,
                                               ShardedTableStatistics, env_rank_world)
    import bench
    import gibbs_bench
    rank, world, local = env_rank_world()
    saved = os.dup(1)
    os.dup2(2, 1)  # NCCL's banner goes to stderr; stdout carries the JSON line only
    if a.all is not None:
        comm = Communicator.init_all([int(x) for x in a.all.split(",")])
        dev0 = comm.devices[0]
    else:
        comm = Communicator.from_env()
        dev0 = local
    bn.set_default_device(dev0)
    comm.barrier()
    G = comm.world
    per_gpu = int(a.particles or (2e7 if a.quick else 1e9))

    def check(net, lay, query, evidence):
        im = bn.LikelihoodWeightingInference(nsamples=4096, seed=1, device=dev0)
        im.raw_counts(bn.InferenceState(net, query, evidence))
        return im.last_stats[0] > 0

    net, lay, query, evidence = bench.build_case(bn, check=check)
    bn.set_lw_mode("specialized", False)
    bn.set_gibbs_mode("specialized")
    res = {"world": G, "n_local": comm.n_local, "mode": "one process, all GPUs" if a.all is not None else "process per GPU"}

    def timed(job, reps):
        for w in range(2):
            job.seed = 100 + w
            job.step()
        comm.barrier()
        t0 = time.perf_counter()
        for r in range(reps):
            job.seed = r
            job.step()
        comm.barrier()
        return (time.perf_counter() - t0) / reps

    # ---- LW: the whole job through the comm net vs the same particle range on one GPU
    total = per_gpu * G
    job = ShardedLikelihoodWeighting(net, query, evidence, total, seed=0, comm=comm)
    dt = timed(job, 3)
    got = job.counts()
    one = ShardedLikelihoodWeighting(net, query, evidence, total if a.quick else per_gpu, seed=2, device=dev0)
    job.seed = 2
    job.n_total = one.n_total
    job.step()
    got = job.counts()
    one.step()
    ref = one.counts()
    res["lw"] = {"particles": total, "ms": dt * 1e3, "particles_per_s": total / dt}
    res["lw_equal_single_gpu_1e12"] = bool(np.allclose(got, ref, rtol=1e-12, atol=0))
    # ---- Gibbs C4
    gnet, glay, gq, gev = gibbs_bench.build_case(bn)
    nch = 4096 if a.quick else 65536
    gj = ShardedGibbs(gnet, gq, gev, nch, 2000, 1000, 5, seed=0, comm=comm)
    gdt = timed(gj, 3)
    gj.seed = 2
    gj.step()
    gg = gj.counts()
    g1 = ShardedGibbs(gnet, gq, gev, nch, 2000, 1000, 5, seed=2, device=dev0)
    g1.step()
    res["gibbs_c4"] = {"n_chains": nch, "ms": gdt * 1e3, "node_updates_per_s": nch * 2000 / gdt}
    res["gibbs_counts_identical"] = bool(np.array_equal(gg, g1.counts()))
    # ---- row-sharded table statistics (process-per-GPU only: one table shard per process)
    if comm.n_local == 1:
        rows = (1 << 18) if a.quick else (1 << 24)
        sh = ShardedTableStatistics(net, rows * G, seed=3, comm=comm)
        sh.step()
        comm.barrier()
        t0 = time.perf_counter()
        sh.step()
        comm.barrier()
        tdt = time.perf_counter() - t0
        whole = ShardedTableStatistics(net, rows * G if a.quick else rows, seed=3, device=dev0)
        if not a.quick:  # compare on a prefix-sized job: re-run the sharded job at that size
            sh = ShardedTableStatistics(net, rows, seed=3, comm=comm)
            sh.step()
        whole.step()
        res["table"] = {"rows": rows * G, "ms": tdt * 1e3, "rows_per_s": rows * G / tdt}
        res["table_counts_identical"] = bool(np.array_equal(sh.counts_host(), whole.counts_host()))
        res["table_logpdf_identical"] = bool(sh.logpdf() == whole.logpdf())
    else:
        res["table_counts_identical"] = True
    comm.barrier()
    os.dup2(saved, 1)
    if comm.first_rank == 0:
        print(json.dumps(res), flush=True)
    ok = res["lw_equal_single_gpu_1e12"] and res["gibbs_counts_identical"] and res["table_counts_identical"]
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
