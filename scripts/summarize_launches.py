#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
of ONE forward step: per-kernel launch counts, time and DRAM traffic, each kernel's share of the step, and the JSON
bench.py reads for `roofline.traffic` (profiles/r01_conv_dram_traffic.json).

    python scripts/summarize_launches.py gpurun_out/step_launches.csv [--json profiles/r01_conv_dram_traffic.json]
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    per_id = defaultdict(dict)
    for r in rd:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_id[r["ID"]]["name"] = r["Kernel Name"]
        per_id[r["ID"]][r["Metric Name"]] = v * scale
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per_id.values():
        name = d["name"]
        m = re.search(r"(cft_[a-z0-9_]+kernel|[a-z0-9_]+_kernel)", name)
        key = m.group(1) if m else name[:60]
        a = agg[key]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    total_us = sum(a[1] for a in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'us':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {a[0]:5d} {a[1]:10.1f} {a[1] / total_us * 100:6.1f}% {a[2] / 1e6:11.1f} {a[3] / 1e6:11.1f}")
    print(f"{'total':44s} {sum(a[0] for a in agg.values()):5d} {total_us:10.1f}")
    conv = agg.get("cft_conv_tcgen05_kernel")
    if out_json and conv:
        json.dump({
            "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                      "one eager step (bench.py --ncu-range), batch 32 @640, serialized launches (cold-cache durations: "
                      "compare shares, not absolutes); summarised by scripts/summarize_launches.py",
            "launches_per_step": conv[0], "ncu_time_us_per_step": conv[1],
            "dram_bytes_read_per_step": conv[2], "dram_bytes_write_per_step": conv[3],
            "dram_bytes_per_step": conv[2] + conv[3], "share_of_step_time": conv[1] / total_us,
        }, open(out_json, "w"), indent=1)
        print("wrote", out_json)


if __name__ == "__main__":
    main()
