#!/usr/bin/env python3
"""Writes the bench workload's upstream bodies as [u32 little-endian length][bytes]... for baseline/go/stream_bench_test.go.
usage: python tools/dump_workload.py C4 65536 /tmp/c4.bin"""
import struct
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inference_gateway_b200 import synth  # noqa: E402

name, n, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
streams, _ = synth.make_config(name, n_streams=n)
with open(path, "wb") as f:
    for body, _, _ in streams:
        f.write(struct.pack("<I", len(body)))
        f.write(body)
print(f"{len(streams)} streams, {sum(len(b) for b, _, _ in streams)} bytes -> {path}")
