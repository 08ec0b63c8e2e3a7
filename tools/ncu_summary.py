#!/usr/bin/env python
"""Summarise an .ncu-rep (captured with --set full) into a markdown table: duration, DRAM traffic and %, tensor-pipe %,
registers, grid.  Runs here (no GPU):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--layers opbench.json] [--static profiles/ncu_conv_static.json]

--layers: the opbench --once JSON of the same run (row i = launch i): adds layer names and algorithmic bytes.
--static: also write the committed numbers bench.py reports as `roofline.traffic` / `tensor_pipe_active_pct_ncu`
          (mean DRAM bytes per conv launch, time-weighted tensor-pipe activity) together with the SHA-256 of the kernel
          sources they were measured on, so that bench.py can flag them stale when the kernels change.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

M_TIME = "gpu__time_duration.sum"
M_RD, M_WR = "dram__bytes_read.sum", "dram__bytes_write.sum"
M_TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
M_REGS, M_GRID, M_BLOCK = "launch__registers_per_thread", "launch__grid_size", "launch__block_size"


def to_float(v, unit, want):
    """ncu prints a unit row; normalise to ms / bytes."""
    x = float(v.replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit) if want == "ms" else \
        {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return x * (scale if scale else 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--layers")
    ap.add_argument("--static")
    ap.add_argument("--source", default=None, help="text stored as the capture's provenance in --static")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    names = None
    if args.layers:
        names = [r["kernel"] for r in json.load(open(args.layers))["rows"]]
    recs = []
    for r in rows[2:]:
        if len(r) <= idx["Kernel Name"]:
            continue
        rec = {"kernel": r[idx["Kernel Name"]].split("(")[0][-60:],
               "ms": to_float(r[idx[M_TIME]], units[idx[M_TIME]], "ms"),
               "dram": to_float(r[idx[M_RD]], units[idx[M_RD]], "B") + to_float(r[idx[M_WR]], units[idx[M_WR]], "B"),
               "tensor": float(r[idx[M_TENSOR]]) if M_TENSOR in idx else float("nan"),
               "regs": r[idx[M_REGS]], "grid": r[idx[M_GRID]], "block": r[idx[M_BLOCK]]}
        recs.append(rec)
    print("| # | layer | kernel instance | time ms | DRAM read+write GB | tensor % | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|")
    for i, r in enumerate(recs):
        layer = names[i] if names and i < len(names) else ""
        print(f"| {i} | {layer} | `{r['kernel']}` | {r['ms']:.3f} | {r['dram'] / 1e9:.3f} | {r['tensor']:.1f} | {r['regs']} | {r['grid']} x {r['block']} |")
    tot_ms = sum(r["ms"] for r in recs)
    tot_dram = sum(r["dram"] for r in recs)
    tw = sum(r["ms"] * r["tensor"] for r in recs) / tot_ms
    print(f"| | **all {len(recs)}** | | **{tot_ms:.2f}** | **{tot_dram / 1e9:.2f}** ({tot_dram / len(recs) / 1e9:.4f} per launch) | **{tw:.1f}** (time-weighted) | | |")
    if args.static:
        from bench import kernel_source_hash
        json.dump({"dram_bytes_per_launch": tot_dram / len(recs), "tensor_pipe_active_pct": tw, "launches": len(recs),
                   "ncu_ms_total": tot_ms, "kernel_source_sha16": kernel_source_hash(),
                   "source": args.source or os.path.basename(args.report)}, open(args.static, "w"), indent=1)
        print(f"\nwrote {args.static}")


if __name__ == "__main__":
    main()
