#!/usr/bin/env python3
"""Device-resident kernel time of ONE steady-state tick: every connection delivers a single SSE event (one short segment per
connection), the shape of real streaming traffic. Complements bench.py (whole streams per micro-batch) and latency_bench.py
(wall time through the C ABI). Prints one JSON object."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--tick", type=int, default=2)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--flags", type=int, default=0, help="sse_config.flags (4 tile kernel, 16 skeleton templates)")
    args = ap.parse_args()
    import torch
    from inference_gateway_b200 import SseEngine, synth
    streams, mode = synth.make_config("C4", n_streams=args.streams)
    evs = [[e + b"\n\n" for e in b.split(b"\n\n") if e] for b, _, _ in streams]
    items = [(c, mode, ev[args.tick]) for c, ev in enumerate(evs) if args.tick < len(ev)]
    nbytes = sum(len(d) for _, _, d in items)
    eng = SseEngine(device=0, max_conns=args.streams, bytes_per_batch=nbytes + 16 * len(items) + 64, n_slots=1, carry_slot_bytes=16384, flags=args.flags)
    stream = torch.cuda.Stream()
    slot, arena, segs = eng.acquire()
    n, nb = eng.fill(arena, segs, items)
    segs["provider"][:n] = np.arange(n) % 4
    eng.upload(slot, n, nb, stream.cuda_stream)
    torch.cuda.synchronize()

    def step():
        eng.reset_all(stream.cuda_stream)
        eng.launch(slot, n, stream.cuda_stream)

    with torch.cuda.stream(stream):
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res = eng.download(slot, stream.cuda_stream)
    print(json.dumps({"what": "one steady-state tick, device resident (reset + kernels)", "segments": n, "bytes": nbytes,
                      "frames": int(res.raw.n_frames), "ms_per_tick": ms, "chunks_per_s": int(res.raw.n_frames) / ms * 1e3}))
    eng.close()


if __name__ == "__main__":
    main()
