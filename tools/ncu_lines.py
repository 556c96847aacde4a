#!/usr/bin/env python3
"""Ranks source lines of an .ncu-rep by executed warp instructions (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None
data = []
ci = ti = si = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        ci, ti, si = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples")
        continue
    if len(r) > 2 and r[2] == "-" and r[0].isdigit():
        try:
            data.append((int(r[ci]), int(r[ti]), int(r[si]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
tots = sum(d[2] for d in data) or 1
print(f"total warp instructions {tot}, samples {tots}")
byfile = {}
for n, t, s_, f, l, src in data:
    byfile[f] = byfile.get(f, 0) + n
print({k: f"{v / tot * 100:.1f}%" for k, v in byfile.items()})
data.sort(reverse=True)
for n, t, s_, f, l, src in data[:top]:
    print(f"{n / tot * 100:5.1f}% lanes {t / max(n, 1):4.1f} smp {s_ / tots * 100:4.1f}% {f}:{l}: {src}")
