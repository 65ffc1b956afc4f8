#!/usr/bin/env python
"""LoopyBelief across GPUs: `per_gpu` (query, evidence) pairs PER GPU on the C2 net (weak scaling), pairs sharded by
index, no data-path collective; the beliefs are all-gathered once outside the timed region for the parity check.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/loopy_scaling.py [per_gpu] [iters]
Prints one JSON line on rank 0: instance-iterations/s (max over ranks of the device time) and whether the gathered
beliefs equal the ones rank 0 computes alone for the first and last shard."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    import torch
    import torch.distributed as dist
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200.distributed import ShardedLoopy, env_rank_world, shard_range
    import loopy_bench
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
    per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    n_total = per_gpu * world
    net, lay, ev, q = loopy_bench.build_case(bn, n_total)
    queries = [lay.names[int(i)] for i in q]
    evidences = [{lay.names[i]: int(e[i]) + 1 for i in np.nonzero(e >= 0)[0]} for e in ev]
    sl = ShardedLoopy(net, queries, evidences, nsamples=iters, tol=-1.0, device=local)
    for _ in range(3):
        sl.launch()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sl.launch()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    allb = sl.gather().cpu().numpy()
    if rank == 0:
        ok = True
        for g in sorted({0, world - 1}):
            one = ShardedLoopy(net, queries, evidences, nsamples=iters, tol=-1.0, device=local)
            first, count = shard_range(n_total, g, world)
            sub = ShardedLoopy.__new__(ShardedLoopy)
            sub.__dict__.update(one.__dict__)
            sub.ev_d = torch.from_numpy(ev[first:first + count]).to(f"cuda:{local}")
            sub.q_d = torch.from_numpy(q[first:first + count]).to(f"cuda:{local}")
            sub.count = count
            sub.out = torch.zeros((count, one.stride), dtype=torch.float64, device=f"cuda:{local}")
            sub.iters = torch.zeros(count, dtype=torch.int32, device=f"cuda:{local}")
            ok = ok and np.array_equal(sub.launch().cpu().numpy(), allb[first:first + count], equal_nan=True)
        t = float(ms.item()) * 1e-3
        print(json.dumps({"config": f"C2 net, LoopyBelief {iters} iterations (tol < 0), {per_gpu} (query, 10 evidence) pairs per GPU",
                          "n_gpus": world, "pairs": n_total, "ms_per_launch": t * 1e3,
                          "instance_iterations_per_s": n_total * iters / t, "scaling": "weak",
                          "gathered_beliefs_equal_single_gpu": bool(ok)}), flush=True)
    if world > 1:
        dist.destroy_process_group()
