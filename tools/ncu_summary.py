#!/usr/bin/env python
"""Summarise an .ncu-rep (captured with --set full) into a markdown table: duration, DRAM traffic and %,
tensor-pipe %, occupancy, registers.  Runs here (no GPU): `python tools/ncu_summary.py gpurun_out/prof.ncu-rep`."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(header)}
    name_i = idx.get("Kernel Name")
    cols = [c for c in WANT if c in idx]
    print("| kernel | " + " | ".join(c.replace("avg.pct_of_peak_sustained_", "%") for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        if len(r) <= name_i:
            continue
        nm = r[name_i]
        nm = nm[:60]
        vals = []
        for c in cols:
            v = r[idx[c]]
            u = units[idx[c]]
            vals.append(f"{v} {u}".strip())
        print(f"| {nm} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
