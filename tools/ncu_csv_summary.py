#!/usr/bin/env python
"""Markdown summary of an `ncu --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,dram__bytes_read.sum,
dram__bytes_write.sum` log: per-kernel totals and, for kernels matching --detail, one row per launch.

    python tools/ncu_csv_summary.py gpurun_out/encoder.csv --title "..." --command "..." --detail modconv3x3_tcr > profiles/x.md
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--command", default="")
    ap.add_argument("--detail", default=None, help="regex: kernels listed launch by launch")
    args = ap.parse_args()
    lines = [l for l in open(args.csv) if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = launches.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""), "grid": row.get("Grid Size", ""),
                                            "ms": 0.0, "tensor": 0.0, "bytes": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        u, m = row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["ms"] = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        elif m.startswith("dram__bytes"):
            d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        elif "tensor" in m:
            d["tensor"] = v
    agg = collections.OrderedDict()
    for d in launches.values():
        key = re.sub(r"<.*", "", d["name"])
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["ms"]
        a[2] += d["tensor"] * d["ms"]
        a[3] += d["bytes"]
    print(f"# {args.title}\n")
    if args.command:
        print(f"Command: {args.command}\n")
    print("| kernel | launches | total ms | time-weighted tensor % | DRAM GB |\n|---|---|---|---|---|")
    for k, (n, ms, tw, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ms:.3f} | {tw / ms if ms else 0:.1f} | {by / 1e9:.2f} |")
    if args.detail:
        print(f"\nPer launch of `{args.detail}` (call order):\n\n| # | kernel instance | grid | ms | tensor % | DRAM GB |\n|---|---|---|---|---|---|")
        for i, d in launches.items():
            if re.search(args.detail, d["name"]):
                print(f"| {i} | `{d['name']}` | {d['grid']} | {d['ms']:.3f} | {d['tensor']:.1f} | {d['bytes'] / 1e9:.3f} |")


if __name__ == "__main__":
    main()
