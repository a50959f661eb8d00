#!/usr/bin/env python
"""Per-kernel totals and shares from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).

    python tools/launch_summary.py gpurun_out/<tag>/launches.csv "<title>" > profiles/<tag>_launches_summary.md
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    lines = [ln for ln in open(path, errors="replace") if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    total, count = defaultdict(float), defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        value = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = value / 1e3 if unit in ("ns", "nsecond") else value * ({"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1))
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("opv::", "").strip()
        total[name] += us
        count[name] += 1
    grand = sum(total.values())
    print(f"# {title}\n\n| kernel | launches | total us | share |\n|---|---|---|---|")
    for name, us in sorted(total.items(), key=lambda kv: -kv[1]):
        print(f"| `{name}` | {count[name]} | {us:.1f} | {100 * us / grand:.1f}% |")
    print(f"\n{sum(count.values())} launches, {grand / 1e3:.2f} ms in total")


if __name__ == "__main__":
    main()
