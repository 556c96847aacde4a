#!/usr/bin/env python3
"""Compares the dump of baseline/go/stream_bench_test.go::TestDumpParity (the UNMODIFIED reference run wherever Go exists)
with this repository's oracle on the same workload file: per stream, frame counts and SHA-256 of the emitted bytes in mode P
and mode R. A clean run is the byte-level pin DESIGN.md section 6 lacks.
usage: python tools/check_go_dump.py /tmp/c4.bin /tmp/c4_go.jsonl"""
import hashlib
import json
import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orc  # noqa: E402  (checker, not product)

raw = open(sys.argv[1], "rb").read()
streams, off = [], 0
while off + 4 <= len(raw):
    n = struct.unpack_from("<I", raw, off)[0]
    streams.append(raw[off + 4:off + 4 + n])
    off += 4 + n
bad = 0
for line in open(sys.argv[2]):
    d = json.loads(line)
    body = streams[d["stream"]]
    p = orc.passthrough(body)
    r = orc.reframe(body)
    rf = [ln.out for ln in r.lines if ln.kind == orc.L_EMITTED]
    exp = (len(p.lines), hashlib.sha256(p.out).hexdigest(), len(rf), hashlib.sha256(b"".join(rf)).hexdigest())
    got = (d["p_frames"], d["p_sha256"], d["r_frames"], d["r_sha256"])
    if exp != got:
        bad += 1
        print(f"stream {d['stream']}: go {got} != oracle {exp}")
print(f"{len(streams)} streams, {bad} mismatches")
sys.exit(1 if bad else 0)
