#!/usr/bin/env python
"""The learn-from-samples loop across GPUs: rand(bn, n) -> statistics -> logpdf with the table's ROWS sharded by global
row index (weak scaling: rows per GPU fixed), one all-reduce of the int64 counts.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/table_scaling.py [rows_per_gpu]
Prints one JSON line on rank 0: rows/s of sample -> count -> all-reduce (max over ranks of the device time), whether the
all-reduced counts equal the counts one GPU produces for the whole row range (rank 0 recomputes them, when they fit), and
the log-likelihood from the all-reduced counts next to the single-GPU one."""
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
    from bayesnets_jl_b200.distributed import ShardedTableStatistics, env_rank_world
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
    rows_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
    n_rows = rows_per_gpu * world
    net = bn.rand_discrete_bn(100, 3, 4, seed=0)  # the C2 net
    bn.set_lw_mode("specialized", True)  # the network-specialised table sampler from the first call (as bench.py does)
    sh = ShardedTableStatistics(net, n_rows, seed=0, device=local)
    for w in range(3):
        sh.seed = 100 + w
        sh.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        sh.seed = r
        sh.step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    counts = sh.counts.cpu().numpy().copy()
    logl = sh.logpdf()
    if rank == 0:
        line = {"config": "C2 net (100 nodes), rand(bn, n) -> statistics -> logpdf, rows sharded by global row index",
                "n_gpus": world, "rows_per_gpu": rows_per_gpu, "rows": n_rows, "scaling": "weak", "ms_per_step": float(ms.item()),
                "rows_per_s": n_rows / (float(ms.item()) * 1e-3), "logpdf": logl, "logpdf_per_row": logl / n_rows}
        if n_rows * 100 <= 8 << 30:  # the whole table fits beside rank 0's shard
            one = ShardedTableStatistics(net, n_rows, seed=reps - 1, device=local)
            one.first, one.count = 0, n_rows
            one.table = torch.empty((one.table.shape[0], n_rows), dtype=torch.uint8, device=f"cuda:{local}")
            one.sample()
            one.launch()
            line["counts_equal_single_gpu"] = bool(np.array_equal(counts, one.counts.cpu().numpy()))
            line["logpdf_single_gpu"] = one.logpdf()
            line["logpdf_bits_equal"] = bool(line["logpdf_single_gpu"] == logl)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
