#!/usr/bin/env python
"""BASELINE config C4 across GPUs: 65 536 chains sharded by global chain index, one all-reduce of the counts.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/gibbs_scaling.py
Prints one JSON line on rank 0: node-updates/s (max over ranks of the device time), and whether the all-reduced integer
counts equal the counts the same ranks produce for the full chain range on one GPU (rank 0 recomputes them)."""
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
    from bayesnets_jl_b200.distributed import ShardedGibbs, env_rank_world
    import gibbs_bench
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
    n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    net, lay, query, evidence = gibbs_bench.build_case(bn)
    bn.set_gibbs_mode("specialized")
    sg = ShardedGibbs(net, query, evidence, n_chains, 2000, 1000, 5, seed=0, device=local)
    for w in range(3):
        sg.seed = 100 + w
        sg.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        sg.seed = r
        sg.step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    counts = sg.out.cpu().numpy().copy()
    if rank == 0:
        one = ShardedGibbs(net, query, evidence, n_chains, 2000, 1000, 5, seed=reps - 1, device=local)
        one.first, one.count = 0, n_chains  # the whole chain range on this GPU
        full = one.launch().cpu().numpy()
        print(json.dumps({"config": "C4: 200-node net, 20 evidence, GibbsSamplingNodewise(2000, 1000, 5)", "n_gpus": world,
                          "n_chains": n_chains, "ms_per_call": float(ms.item()),
                          "node_updates_per_s": n_chains * 2000 / (float(ms.item()) * 1e-3),
                          "counts_equal_single_gpu": bool(np.array_equal(counts, full)),
                          "posterior": [round(float(x), 5) for x in counts / counts.sum()]}), flush=True)
    if world > 1:
        dist.destroy_process_group()
