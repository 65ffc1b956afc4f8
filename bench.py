#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: likelihood-weighting samples/s (+ factor GB/s).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (CPU arm: the oracle restatement on host cores)

Workload (config C2 of BASELINE.json): synthetic rand_discrete_bn(100, max_parents=3, max_states=4)
net seed 0, 10 evidence nodes with P(e) > 0, 2 query nodes; one STEP = one LW inference pass of 1e9 particles
PER GPU (weak scaling: N GPUs -> N x 1e9 particles, sharded by global particle id, one all-reduce of the
weighted-count tensor).  `value` = whole-job particles/s with the network resident in HBM; `e2e` = the same
through the public infer() call from host buffers (network upload + result download inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# from the committed ncu capture of the headline kernel (profiles/r1c_lw_spec_ncu_full.txt): 5.179e9 warp
# instructions for 1e8 particles (3.125e6 warps), DRAM bytes per launch, pipe utilisation
NCU_LW = {"inst_per_warp_particle": 5178894080 / (1e8 / 32), "dram_bytes_per_launch": 45312,
          "fmaheavy_pct": 80.2, "alu_pct": 61.7, "issue_active_pct": 61.7}
METRIC = "lw_samples_per_s"
UNIT = "samples/s"


# ------------------------------------------------------------------------------------------ workload
def build_case(bn, n_nodes=100, net_seed=0, n_ev=10, n_q=2, check=None):
    """The C2 case.  `check(query, evidence) -> bool` rejects evidence with P(e) = 0."""
    net = bn.rand_discrete_bn(n_nodes, 3, 4, seed=net_seed)
    lay = bn.flatten(net)
    rng = np.random.default_rng(1000 + net_seed)
    while True:
        picks = [str(x) for x in rng.choice(lay.names, size=n_q + n_ev, replace=False)]
        query, evn = picks[:n_q], picks[n_q:]
        evidence = {e: int(rng.integers(1, net.ncategories(e) + 1)) for e in evn}
        if check is None or check(net, lay, query, evidence):
            return net, lay, query, evidence


def algorithmic_bytes_per_particle(lay) -> float:
    """LW particle = sum_i 8*card_i (the CPT row touched per node, fp64-equivalent)."""
    return float(8 * int(lay.card.sum()))


def live_node_count(lay, query, evidence) -> int:
    """Nodes that can influence the query or the weight: ancestors of query + evidence through free parents."""
    ev = {lay.index[e] for e in evidence}
    stack, live = [lay.index[q] for q in query] + sorted(ev), set()
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        stack.extend(p for p in lay.parents(i) if p not in ev)
    return len(live)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_lw(lay, query, evidence, n, seed, threads=None):
    """Oracle LW (reference restatement, likelihood.jl:34-54 loop) on host cores -> (seconds, threads)."""
    from oracle import capi
    from oracle.factors_np import FlatNet
    capi.build()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core it is allowed to run on
    capi.set_num_threads(threads or len(os.sched_getaffinity(0)))
    on = FlatNet(list(lay.names), lay.card, lay.par_ptr, lay.par_idx, lay.cpt_off, lay.cpt, dict(lay.index))
    evv, qi = lay.evidence_vector(evidence), lay.query_indices(query)
    t0 = time.perf_counter()
    c, sw, sw2 = capi.lw_infer(on, evv, qi, n, seed=seed, mode=capi.REFERENCE_SCAN)
    return time.perf_counter() - t0, capi.num_threads(), c


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia itself cannot run here) with all host
    threads, same config/metric; each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bayesnets_jl_b200 as bn
    net, lay, query, evidence = build_case(bn)
    dt, threads, _ = cpu_lw(lay, query, evidence, 200_000, 1)
    per_step = int(max(200_000, min(5e7, 200_000 / dt * 4.0)))  # ~4 s of CPU work per step
    for w in range(args.warmup):
        cpu_lw(lay, query, evidence, per_step // 4, 100 + w)
    t = 0.0
    for k in range(args.steps):
        dt, threads, _ = cpu_lw(lay, query, evidence, per_step, 200 + k)
        t += dt
    value = per_step * args.steps / t
    sample = f"{per_step} particles/step x {args.steps} steps of the C2 workload (full workload: 1e9/step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(lay, query, evidence, per_step, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = C/OpenMP restatement of src/Inference/likelihood.jl:34-54 (oracle/bn_oracle.c); the "
                "Julia package itself cannot run in this image (no julia, deps un-vendored)",
    }
    print(json.dumps(line), flush=True)


def workload_config(lay, query, evidence, per_gpu, n_gpus):
    return {
        "workload": "C2: LW infer on rand_discrete_bn(100,3,4) net_seed=0, 10 evidence, 2 query",
        "n_nodes": int(lay.n), "cpt_entries": int(lay.cpt.size), "n_evidence": len(evidence),
        "query_cells": int(np.prod([lay.card[lay.index[q]] for q in query])),
        "particles_per_step_per_gpu": int(per_gpu), "particles_per_step": int(per_gpu) * n_gpus,
        "parallelism": f"particles sharded over {n_gpus} GPU(s) by global Philox counter + 1 all-reduce of counts",
        "l2": "n/a - no HBM-resident operands (tables ~20 KB are TMA-staged to smem per launch); "
              "fresh seed every step, nothing cached",
    }


# ------------------------------------------------------------------------------------------ factor sweep
def factor_sweep(bn, torch, device, k=24, iters=5):
    """Factor `*` / sum / normalize / fused eliminate GB/s at 2^k entries (C5 shapes) through the C ABI on
    device-resident handles: 24 ops back to back per CUDA-event pair on the launching stream, rotating over 3 distinct
    operand sets (working set ~1 GB >> 126 MB L2, so nothing is served from cache) -- tools/factor_bench.py."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import factor_bench
    hbm, src = peaks()
    res = []
    for r in factor_bench.run(bn, torch, device=device, k=k):
        res.append({"case": r["case"], "bytes": r["bytes"], "ms": r["batch_ms"], "GB/s": r["batch_GBs"],
                    "frac": round(r["batch_GBs"] / hbm, 3), "single_call_ms": r["call_ms"]})
    # the ceiling at THIS size: a plain device copy of 2^k doubles timed the same way (MEASURED_PEAKS used 2 GiB)
    n = 1 << k
    srcs = [torch.rand(n, dtype=torch.float64, device=f"cuda:{device}") for _ in range(3)]
    dsts = [torch.empty(n, dtype=torch.float64, device=f"cuda:{device}") for _ in range(3)]
    for i in range(6):
        dsts[i % 3].copy_(srcs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(24):
        dsts[i % 3].copy_(srcs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    copy_gbs = 16.0 * n * 24 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return res, round(copy_gbs, 1)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import bayesnets_jl_b200 as bn
    from bayesnets_jl_b200 import _lib
    from bayesnets_jl_b200.distributed import ShardedLikelihoodWeighting, env_rank_world

    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL prints its version banner (and any NCCL_DEBUG output) to fd 1
        # while the communicator is created, so fd 1 points at stderr until the first collective has run
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
    if args.gpus != world and rank == 0:
        print(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    torch.cuda.set_device(local)
    bn.set_default_device(local)
    per_gpu = int(args.samples)

    def check(net, lay, query, evidence):  # P(e) > 0 via a short LW run on the device
        im = bn.LikelihoodWeightingInference(nsamples=4096, seed=1, device=local)
        im.raw_counts(bn.InferenceState(net, query, evidence))
        return im.last_stats[0] > 0

    net, lay, query, evidence = build_case(bn, check=check)
    # HEADLINE kernel: the network-specialised LW kernel with EVERY node sampled (prune off).  The product default
    # additionally skips barren nodes (bit-identical counts, ~2.3x fewer node-samples on this case); that and the
    # generic interpreter are reported under "variants", never as `value`.
    headline = {"specialized_all": ("specialized", False), "specialized_pruned": ("specialized", True),
                "generic": ("generic", True)}[args.lw_kernel]
    bn.set_lw_mode(*headline)
    slw = ShardedLikelihoodWeighting(net, query, evidence, per_gpu * world, seed=0, device=local)
    stream = torch.cuda.current_stream(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(args.warmup, 3)):
        slw.seed = 1000 + w
        slw.step()
    barrier()
    # ---- timed region: exactly K steps, events on the launching stream
    launches0 = bn.launch_count()
    kern_ev = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        for k in range(args.steps):
            slw.seed = k
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            slw.launch()
            b.record(stream)
            kern_ev.append((a, b))
            if world > 1:
                dist.all_reduce(slw.out)
        e1.record(stream)
        barrier()
    launches = bn.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))
    posterior = slw.posterior()
    n_eff = float(slw.out[slw.ncell] ** 2 / slw.out[slw.ncell + 1])

    # ---- e2e through the public API with host buffers (network flattened + uploaded every step)
    e2e_steps = max(1, min(args.steps, 3))
    h2d = int(lay.card.nbytes + lay.par_ptr.nbytes + lay.par_idx.nbytes + lay.cpt_off.nbytes + lay.cpt.nbytes + 24 * 1024)
    d2h = int((slw.ncell + 2) * 8 + slw.ncell * 8)
    first, count = slw.first, slw.count
    e2e_step_ms = []

    def e2e_step(k):
        ts = time.perf_counter()
        net._device_cache = None  # force flatten + bn_net_create (H2D) inside the timed region
        im = bn.LikelihoodWeightingInference(nsamples=count, seed=500 + k, device=local, first_particle=first)
        if world > 1:
            c = torch.from_numpy(im.raw_counts(bn.InferenceState(net, query, evidence))).to(f"cuda:{local}")
            dist.all_reduce(c)
            c = c.cpu().numpy()
        else:
            c = bn.infer(im, net, query, evidence=evidence).potential
        e2e_step_ms.append(round(1e3 * (time.perf_counter() - ts), 3))
        return c

    e2e_step(-1)  # one untimed warm-up of the host-buffer entry point (first-call allocations)
    e2e_step_ms.clear()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = per_gpu * world * e2e_steps / float(e2e_s.item())

    # ---- other kernel variants on the same workload (short, device-timed; not the headline)
    variants = {}
    if world == 1:
        for name, (mode, prune) in {"specialized_all": ("specialized", False), "specialized_pruned": ("specialized", True),
                                    "generic": ("generic", True)}.items():
            if name == args.lw_kernel:
                continue
            bn.set_lw_mode(mode, prune)
            for w in range(2):
                slw.seed = 2000 + w
                slw.launch()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            for k in range(2):
                slw.seed = 3000 + k
                slw.launch()
            b.record(stream)
            torch.cuda.synchronize()
            variants[name] = {"value": per_gpu * 2 / (a.elapsed_time(b) * 1e-3), "unit": UNIT,
                              "posterior": [round(float(x), 6) for x in slw.posterior().ravel(order="F")]}
        bn.set_lw_mode(*headline)
    live = live_node_count(lay, query, evidence)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, src = peaks()
    # the bound that actually applies: issue slots of the busiest pipe.  Instructions per warp-particle are a static
    # property of the generated kernel (ncu smsp__inst_executed.sum / warps, profiles/r1c_lw_spec_ncu_full.txt);
    # the live fraction below is that count x warps launched / (SMs x 4 schedulers x clock x measured duration)
    sm_clock = (line_clock := clk.summary()).get("sm_mhz") or 1965.0
    warp_instr = NCU_LW["inst_per_warp_particle"] * per_gpu / 32.0
    pipe = {"bound": "issue slots; busiest pipe = fmaheavy (IMAD.WIDE of Philox4x32-10) then alu",
            "issue_slot_frac_live": warp_instr / (148 * 4 * sm_clock * 1e6 * kern_ms * 1e-3),
            "ncu_fmaheavy_pct": NCU_LW["fmaheavy_pct"], "ncu_alu_pct": NCU_LW["alu_pct"],
            "ncu_issue_active_pct": NCU_LW["issue_active_pct"], "inst_per_node_sample": NCU_LW["inst_per_warp_particle"] / lay.n,
            "source": "profiles/r1c_lw_spec_ncu_full.txt (ncu --set full, same command at 1e8 particles)"}
    if args.lw_kernel != "specialized_all":
        pipe = None
    bpp = algorithmic_bytes_per_particle(lay)
    achieved = bpp * per_gpu / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": per_gpu * world * args.steps / (ms_total * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 thresholds / f64 weights",
        "data": "synthetic", "config": workload_config(lay, query, evidence, per_gpu, world),
        "clocks": line_clock,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "warmup": 1, "step_ms": e2e_step_ms, "api": "infer(LikelihoodWeightingInference(nsamples), bn, query; evidence)"},
        "gpu_launches": int(launches),
        "lw_kernel": args.lw_kernel, "nodes_sampled_per_particle": int(lay.n) if args.lw_kernel != "specialized_pruned" else live,
        "variants": variants, "jit_compiles": int(_lib.lib().bn_lw_spec_compile_count()),
        "roofline": {"kernel": {"generic": "bn::particle_kernel<MODE_INFER>"}.get(args.lw_kernel, "lw_spec (NVRTC, network-specialised)"), "bound": "hbm", "achieved": achieved, "peak": hbm,
                     "unit": "GB/s", "frac": achieved / hbm, "traffic": NCU_LW["dram_bytes_per_launch"], "peak_source": f"of {src}",
                     "pipe_roofline": pipe,
                     "algorithmic_bytes_per_particle": bpp, "kernel_ms": kern_ms,
                     "note": "HBM-EQUIVALENT figure required by SURVEY.md 8(d): the CPT rows a particle touches, "
                             "counted as fp64 bytes.  The tables are shared-memory resident, so this kernel is "
                             "bounded by issue slots, not HBM, and frac > 1 is expected; see profiles/ for "
                             "the ncu issue-slot utilisation."},
        "posterior": [round(float(x), 6) for x in posterior.ravel(order="F")], "n_eff": n_eff,
    }
    if world == 1:
        dt, threads, _ = cpu_lw(lay, query, evidence, 200_000, 1)
        n_cpu = int(max(200_000, min(2e8, 200_000 / dt * 12.0)))  # ~12 s of CPU work
        dt, threads, _ = cpu_lw(lay, query, evidence, n_cpu, 2)
        line["cpu_baseline"] = {"value": n_cpu / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n_cpu} particles of the same net/evidence/query (full step: {per_gpu})"}
        if not args.no_factors:
            fr, copy_gbs = factor_sweep(bn, torch, local, k=args.factor_log2)
            line["factor_roofline"] = {"bound": "hbm", "peak": hbm, "peak_source": f"of {src}", "unit": "GB/s",
                                       "log2_entries": args.factor_log2,
                                       "l2": "operands rotate over 3 sets (~1 GB working set >> 126 MB L2)",
                                       "device_copy_same_size_GBs": copy_gbs, "cases": fr}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=float, default=1e9, help="particles per step per GPU")
    ap.add_argument("--factor-log2", type=int, default=24)
    ap.add_argument("--no-factors", action="store_true")
    ap.add_argument("--lw-kernel", default="specialized_all", choices=["specialized_all", "specialized_pruned", "generic"],
                    help="kernel timed as the headline (default: specialised, every node sampled)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
