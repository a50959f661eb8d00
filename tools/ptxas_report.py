#!/usr/bin/env python
"""Per-kernel registers / spills / smem from the `-Xptxas -v` build log (open_provence_b200/lib/build.log)."""
import re, subprocess, sys
log = open(sys.argv[1] if len(sys.argv) > 1 else "open_provence_b200/lib/build.log").read().splitlines()
name = None
rows = []
for i, ln in enumerate(log):
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        spill = used = ""
    if "spill stores" in ln:
        spill = ln.strip()
    m = re.search(r"Used (\d+) registers", ln)
    if m and name:
        rows.append((name, int(m.group(1)), spill))
        name = None
for n, r, s in rows:
    flag = "" if " 0 bytes spill stores" in s else "   <-- " + s
    print(f"{r:4d}  {n[:110]}{flag}")
