#!/usr/bin/env python3
"""Lists the hottest local-memory (LDL/STL) instructions of an .ncu-rep with a few SASS lines of context."""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, ci, si = [], None, None
for r in rows:
    if r and r[0] == "Address":
        ci, si = r.index("Instructions Executed"), r.index("Source")
        continue
    if ci is not None and len(r) > ci:
        data.append(r)
tot = sum(int(r[ci] or 0) for r in data)
loc = [(int(r[ci] or 0), k) for k, r in enumerate(data) if "LDL" in r[si] or "STL" in r[si]]
loc.sort(reverse=True)
print(f"total warp instructions {tot}, local-memory instructions {sum(c for c, _ in loc)}")
for c, k in loc[:top]:
    print(f"--- {c}")
    for q in data[max(0, k - 3):k + 3]:
        print("   ", q[ci].rjust(9), q[si][:110])
