#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of numbers the roofline needs.

    python tools/ncu_summary.py gpurun_out/<tag>/prof_gemm.ncu-rep [more.ncu-rep ...] > profiles/<name>.md

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and prints one block per captured launch.
"""

from __future__ import annotations

import csv
import io
import subprocess
import sys

METRICS = [
    ("time_us", "gpu__time_duration.sum"),
    ("sm_clock_ghz", "sm__cycles_elapsed.avg.per_second"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_throughput_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_slots_busy_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("xu_pipe_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("registers_per_thread", "launch__registers_per_thread"),
    ("dyn_smem_per_block_KB", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("occupancy_limit_blocks_smem", "launch__occupancy_limit_shared_mem"),
    ("occupancy_limit_blocks_regs", "launch__occupancy_limit_registers"),
    ("local_load_bytes", "smsp__inst_executed_op_local_ld.sum"),
    ("local_store_bytes", "smsp__inst_executed_op_local_st.sum"),
]


def to_float(text: str) -> float | None:
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def summarise(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(header)}
    print(f"## {path}\n")
    for r in rows[2:]:
        print(f"### launch {r[col['ID']]}: `{r[col['Kernel Name']][:140]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        vals = {}
        for label, metric in METRICS:
            if metric in col:
                vals[label] = to_float(r[col[metric]])
                print(f"| {label} (`{metric}`) | {r[col[metric]]} | {units[col[metric]]} |")
        rd, wr = vals.get("dram_read_MB"), vals.get("dram_write_MB")
        if rd is not None and wr is not None:
            unit = units[col["dram__bytes_read.sum"]]
            print(f"| **traffic = read + write** | {rd + wr:.3f} | {unit} |")
        print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarise(p)
