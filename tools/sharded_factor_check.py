#!/usr/bin/env python
"""A factor split over the GPUs of one box (bayesnets.jl_b200/sharded_factors.py): parity against the unsplit
single-GPU computation, then the BASELINE C5 shapes at 2^24 entries PER GPU with the one real exchange step of the
path -- the all-reduce that sums out the split variable -- timed on the device.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_factor_check.py
Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    import torch
    import torch.distributed as dist
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200.distributed import env_rank_world
    from bayesnets_jl_b200.sharded_factors import DeviceBackend, ShardedFactor
    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    bn.set_default_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    be = DeviceBackend()
    nsplit = int(np.log2(world))
    assert 1 << nsplit == world, "world size must be a power of two (binary split variables)"
    res = {"n_gpus": world}

    # ---- parity at 2^18 entries: same operands on every rank, compared with the unsplit ops on this GPU
    rng = np.random.default_rng(5)
    k = 18
    names = [f"sf{i}" for i in range(k)]
    A = rng.random((2,) * k)
    sub = list(range(0, k, 3)) + [k - 1]
    B = rng.random((2,) * len(sub))
    fa, fb = bn.Factor(names, A), bn.Factor([names[i] for i in sub], B)
    split = names[k - nsplit:] if nsplit else []
    sa = ShardedFactor.from_replicated(fa, split, rank, world, be)
    prod = sa * fb
    full = fa * fb
    ok_prod = bool(np.array_equal(prod.gather(), np.transpose(full.potential, [full.dimensions.index(d) for d in prod.dimensions])))
    loc = prod.sum(names[4])
    ref = full.sum(names[4])
    ok_loc = bool(np.array_equal(loc.gather(), np.transpose(ref.potential, [ref.dimensions.index(d) for d in loc.dimensions])))
    if nsplit:
        tot = loc.sum(split)
        ref2 = ref.sum(split)
        got = np.transpose(tot.potential, [tot.dimensions.index(d) for d in ref2.dimensions])
        ok_tot = bool(np.allclose(got, ref2.potential, rtol=1e-12, atol=0))
        exact_tot = bool(np.array_equal(got, ref2.potential))
    else:
        ok_tot = exact_tot = True
    nrm = (sa * fb).normalize_()
    refn = (fa * fb).normalize()
    ok_norm = bool(np.allclose(nrm.gather(), np.transpose(refn.potential, [refn.dimensions.index(d) for d in nrm.dimensions]),
                               rtol=1e-12, atol=0))
    res["parity"] = {"product_bit_exact": ok_prod, "local_sum_bit_exact": ok_loc, "split_sum_1e-12": ok_tot,
                     "split_sum_bit_exact": exact_tot, "normalize_1e-12": ok_norm}
    del fa, fb, sa, prod, full, loc, ref, nrm, refn

    # ---- speed: 2^24 entries per GPU (2^(24 + log2 N) in total), P1-style product, local sum, split sum (all-reduce)
    k = 24 + nsplit
    names = [f"sg{i}" for i in range(k)]
    state = ShardedFactor.rank_state([(names[k - nsplit + j], 2) for j in range(nsplit)], rank)
    lrng = np.random.default_rng(100 + rank)
    la = bn.Factor(names[:24], lrng.random(1 << 24).reshape((2,) * 24, order="F"))
    sa = ShardedFactor(la, [(names[k - nsplit + j], 2) for j in range(nsplit)], rank, world, be)
    fb = bn.Factor([names[i] for i in range(0, 24, 2)] + names[24:], np.random.default_rng(7).random((2,) * (12 + nsplit)))

    def timed(fn, reps=5):
        for _ in range(2):
            r = fn(); del r
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn(); del r
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())
    N = 1 << 24
    ms_prod = timed(lambda: sa * fb)
    prod = sa * fb
    ms_lsum = timed(lambda: prod.sum(names[11]))
    res["product_2^24_per_gpu"] = {"ms": ms_prod, "aggregate_GBs": world * 8 * (2 * N + (1 << 12)) / (ms_prod * 1e-3) / 1e9}
    res["local_sum_2^24_per_gpu"] = {"ms": ms_lsum, "aggregate_GBs": world * 8 * (N + N // 2) / (ms_lsum * 1e-3) / 1e9}
    if nsplit:
        half = prod.sum(names[11])                      # 2^23 entries per GPU
        ms_ar = timed(lambda: half.sum([d for d, _ in half.split]))
        nbytes = 8 * (N // 2)
        res["split_sum_allreduce_2^23_per_gpu"] = {"ms": ms_ar, "bytes_per_gpu": nbytes,
                                                   "algbw_GBs": nbytes / (ms_ar * 1e-3) / 1e9,
                                                   "busbw_GBs": nbytes * 2 * (world - 1) / world / (ms_ar * 1e-3) / 1e9,
                                                   "note": "includes the private copy of the local table the all-reduce works on"}
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()
