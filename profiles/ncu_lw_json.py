#!/usr/bin/env python
"""profiles/r2_lw_spec_ncu.json (what bench.py's `roofline` object reads) from an ncu report of the headline kernel.

    python profiles/ncu_lw_json.py gpurun_out/r2b_lw_spec.ncu-rep 1000000000 > profiles/r2_lw_spec_ncu.json

The report is `ncu --set full --clock-control none -k regex:lw_spec -c 1` of
`python bench.py --steps 2 --warmup 3 --no-configs --no-cold --no-factors` (every node sampled, 1e9 particles per launch).
"""
import csv
import io
import json
import subprocess
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, particles):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def g(k, scale=False):
        v = float(r[hdr.index(k)].replace(",", ""))
        return v * SCALE.get(units[hdr.index(k)], 1.0) if scale else v
    j = {"warp_inst_per_particle": g("smsp__inst_executed.sum") / particles,
         "dram_bytes_per_launch": g("dram__bytes_read.sum", True) + g("dram__bytes_write.sum", True),
         "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "fmaheavy_pct": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
         "alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
         "registers_per_thread": g("launch__registers_per_thread"), "block_size": g("launch__block_size"),
         "grid_size": g("launch__grid_size"), "kernel_ms_under_ncu": g("gpu__time_duration.sum"), "particles": particles,
         "source": "profiles/r2_lw_spec_ncu_full.txt = ncu --set full --clock-control none of `python bench.py --steps 2 "
                   "--warmup 3 --no-configs --no-cold --no-factors` (the shipped lw_spec, every node sampled, 1e9 particles "
                   "per launch); made by profiles/ncu_lw_json.py"}
    print(json.dumps(j, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
