#!/usr/bin/env python3
"""Executed warp instructions of one kernel of an .ncu-rep by source region (function of sse_kernel2.cu / header file) and the
top source lines. usage: ncu_regions.py <rep> [kernel index in the report, default 0] [top lines, default 30]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, seen, cur_block = [], set(), None
cur = None
ci = ti = si = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        if cur_block is None or r[1] in seen:
            cur_block = []; blocks.append(cur_block); seen = set()
        seen.add(r[1]); cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ci, ti, si = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples")
        continue
    if len(r) > 2 and r[2] == "-" and r[0].isdigit():
        try:
            cur_block.append((int(r[ci]), int(r[ti]), int(r[si]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
data = blocks[kidx]
tot = sum(d[0] for d in data) or 1
tots = sum(d[2] for d in data) or 1
print(f"kernel #{kidx} of {len(blocks)}: total warp instructions {tot}, samples {tots}")
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "inference_gateway_b200", "csrc", "sse_kernel2.cu")).read().split("\n")
marks = [(i + 1, s.strip()[:64]) for i, s in enumerate(src) if re.match(r"(template|__device__|__global__|struct)", s) or "sse_decode_kernel(" in s]


def owner(f, line):
    if f != "sse_kernel2.cu":
        return f
    o = None
    for ln, s in marks:
        if ln <= line:
            o = s
        else:
            break
    return o


agg, aggs = collections.Counter(), collections.Counter()
for n, t, s_, f, l, srcl in data:
    agg[owner(f, l)] += n; aggs[owner(f, l)] += s_
for k, v in agg.most_common(14):
    print(f"{v / tot * 100:5.1f}% instr {aggs[k] / tots * 100:5.1f}% samples  {k}")
data.sort(reverse=True)
for n, t, s_, f, l, srcl in data[:top]:
    print(f"{n / tot * 100:5.1f}% lanes {t / max(n, 1):4.1f} smp {s_ / tots * 100:4.1f}% {f}:{l}: {srcl}")
