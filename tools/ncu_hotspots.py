#!/usr/bin/env python
"""Instruction-level stall hot spots of one kernel in an ncu report (works without a GPU).

    python tools/ncu_hotspots.py <report.ncu-rep> <kernel-regex> [launch-index] [top-N]

Prints the SASS instructions with the most warp-stall samples and their dominant stall reasons,
followed by a coarse histogram of the samples over the instruction stream (where the time is)."""
import csv, io, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{idx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
h = rows[1]
ci = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(h) or r[0] == "Address" or not r[ci["# Samples"]].replace(".", "").isdigit():
        continue
    n = float(r[ci["# Samples"]] or 0)
    reasons = sorted(((float(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    data.append((n, k, r[ci["Source"]].strip(), reasons, float(r[ci["Instructions Executed"]] or 0)))
total = sum(d[0] for d in data)
print(f"total samples {total:.0f}, {len(data)} instructions")
for n, k, src, reasons, ex in sorted(data, reverse=True)[:top]:
    rs = " ".join(f"{s}:{v:.0f}" for v, s in reasons if v > 0)
    print(f"{n:6.0f} {100 * n / total:5.1f}%  #{k:4d} x{ex:9.0f}  {src[:70]:70s} {rs}")
print("\nsamples per 100-instruction window:")
for a in range(0, len(data), 100):
    s = sum(d[0] for d in data[a:a + 100])
    print(f"  #{a:4d}-{a + 99:4d}: {100 * s / total:5.1f}%  {data[a][2][:60]}")
