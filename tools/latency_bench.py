#!/usr/bin/env python3
"""Added latency of the GPU path in steady-state streaming (SURVEY.md section 8d "p50/p99 added latency").

Every tick delivers the NEXT SSE event of each of N concurrent streams (one small segment per connection, so carry
state, per-connection FIFO and early termination are exercised across ticks), through the public C ABI:
sse_acquire -> fill pinned staging -> sse_submit -> sse_collect -> sse_release. The added latency of a chunk is the
wall time from "its bytes are in host memory" to "its frame is in host memory" = one submit..collect, measured per tick.
Prints one JSON object.
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--rounds", type=int, default=6, help="how many times the stream population is replayed")
    ap.add_argument("--flags", type=int, default=0, help="sse_config.flags (4 tile kernel, 16 skeleton templates)")
    args = ap.parse_args()
    from inference_gateway_b200 import SseEngine, synth
    streams, mode = synth.make_config("C4", n_streams=args.streams)
    events = [[e + b"\n\n" for e in b.split(b"\n\n") if e] for b, _, _ in streams]
    n_ticks = max(len(e) for e in events)
    tick_bytes = [sum(len(ev[t]) for ev in events if t < len(ev)) for t in range(n_ticks)]
    eng = SseEngine(device=0, max_conns=args.streams, bytes_per_batch=max(tick_bytes) + 64, n_slots=2, carry_slot_bytes=16384, flags=args.flags)
    lat, frames_total, bytes_total = [], 0, 0
    t_start = None
    for r in range(args.rounds + 1):                  # round 0 is warm-up
        eng.reset_all()
        for t in range(n_ticks):
            slot, arena, segs = eng.acquire()
            off = 0
            n = 0
            for c, ev in enumerate(events):           # host-side fill (not timed: stands for the socket reads)
                if t < len(ev):
                    d = ev[t]
                    arena[off:off + len(d)] = np.frombuffer(d, dtype=np.uint8)
                    segs[n] = (c, off, len(d), mode, c % 4, 0)   # provider hint: flavours cycle as in bench.py
                    off = (off + len(d) + 15) & ~15
                    n += 1
            t0 = time.perf_counter()
            eng.submit(slot, n, off)
            res = eng.collect(slot)
            dt = time.perf_counter() - t0
            if r > 0:
                if t_start is None:
                    t_start = t0
                lat.append(dt * 1e3)
                frames_total += int(res.raw.n_frames)
                bytes_total += off
            eng.release(slot)
    lat.sort()
    q = lambda p: lat[min(len(lat) - 1, int(p * len(lat)))]
    print(json.dumps({
        "what": "submit->collect wall time per micro-batch (one SSE event per stream per tick), pinned H2D + kernels + D2H",
        "streams": args.streams, "ticks_measured": len(lat), "events_per_tick_mean": frames_total / max(1, len(lat)),
        "bytes_per_tick_mean": bytes_total / max(1, len(lat)),
        "latency_ms": {"p50": q(0.50), "p90": q(0.90), "p99": q(0.99), "max": lat[-1], "mean": statistics.mean(lat)},
        "chunks_per_s_at_this_batching": frames_total / (sum(lat) / 1e3),
    }))
    eng.close()


if __name__ == "__main__":
    main()
