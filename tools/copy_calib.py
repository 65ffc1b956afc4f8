#!/usr/bin/env python
"""HBM ceiling at the factor sweep's sizes: torch copy_ / sum / scale on 2^k doubles, timed like factor_bench (batch)."""
import sys
import torch
k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
srcs = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(3)]
dsts = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
halfs = [torch.empty(n // 2, dtype=torch.float64, device="cuda") for _ in range(3)]
def timed(fn, reps=12):
    for i in range(6): fn(i % 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i % 3)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, fn, nbytes in [
    ("copy n->n (16n bytes)", lambda i: dsts[i].copy_(srcs[i]), 16 * n),
    ("add halves n->n/2 (12n bytes)", lambda i: torch.add(srcs[i][: n // 2], srcs[i][n // 2:], out=halfs[i]), 12 * n),
    ("sum n->1 (8n bytes)", lambda i: srcs[i].sum(), 8 * n),
    ("scale in place (16n bytes)", lambda i: srcs[i].mul_(1.0000001), 16 * n),
]:
    ms = timed(fn)
    print(f"{name:32s} {ms*1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s")
