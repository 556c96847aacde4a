#!/usr/bin/env python3
"""bench.py -- SSE chunks/sec of the streaming-response hot path on B200 (BASELINE.json metric).

A step = one pass of the hot path over one micro-batch of synthetic SSE input: workload C4 of BASELINE.json
(65,536 concurrent mixed cohere/groq/anthropic/ollama streams incl. tool_calls deltas, ~512 B mean chunk,
mode R = MCP reframe + JSON side-band) per GPU, every stream complete in the batch (11 SSE events).

  value   emitted chunks/s with the batch already resident in HBM (kernel path only, CUDA events)
  e2e     the same through the C ABI a caller uses (sse_submit/sse_collect): pinned H2D + kernel + D2H per step
  roofline  algorithmic bytes of the stream kernel / its measured duration vs the measured HBM peak
  cpu_baseline  the CPU oracle port (oracle/, test infrastructure) timed on a bounded sample on the host cores

Multi-GPU (torchrun, one rank per GPU): connections shard by hash(conn_id) % N, no collective on the data path
(weak scaling: 65,536 streams per GPU). `--impl reference` times the reference's CPU path (the oracle port: the
reference is Go and cannot be built here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_SLOTS = 3      # batches in flight on the end-to-end path: H2D, kernel and D2H of consecutive batches overlap
METRIC = "SSE chunks/sec @ 64k concurrent streams"
UNIT = "chunks/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def build_workload(n_streams: int, shard: int, workload: str, n_content=None):
    from inference_gateway_b200 import synth
    t0 = time.time()
    streams, mode = synth.make_config(workload, n_streams=n_streams, shard=shard, n_content=n_content)
    bodies = [b for b, _, _ in streams]
    n_events = sum(n for _, n, _ in streams)
    log(f"[bench] rank shard {shard}: generated {len(bodies)} streams, {sum(map(len, bodies)) / 1e6:.1f} MB, "
        f"{n_events} SSE events in {time.time() - t0:.1f}s")
    return bodies, (mode if mode is not None else 3), n_events


def fill_slot(eng, arena, segs, bodies, mode):
    off = 0
    lens = np.fromiter((len(b) for b in bodies), dtype=np.int64, count=len(bodies))
    offs = np.zeros(len(bodies), dtype=np.int64)
    aligned = (lens + 15) & ~15
    offs[1:] = np.cumsum(aligned)[:-1]
    total = int(offs[-1] + aligned[-1]) if len(bodies) else 0
    for b, o in zip(bodies, offs):
        arena[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    segs["conn"][:len(bodies)] = np.arange(len(bodies), dtype=np.uint32)
    segs["in_off"][:len(bodies)] = offs.astype(np.uint32)
    segs["in_len"][:len(bodies)] = lens.astype(np.uint32)
    segs["mode"][:len(bodies)] = mode
    segs["provider"][:len(bodies)] = np.arange(len(bodies), dtype=np.uint8) % 4   # the connection's provider (flavours cycle cohere/groq/anthropic/ollama)
    segs["reserved"][:len(bodies)] = 0
    return len(bodies), total, int(lens.sum())


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.p = None
        self.idx = gpu_index
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(bodies, mode, threads: int, target_s: float = 3.0):
    """Times the oracle port (CPU restatement of the reference path) on a bounded sample of the workload."""
    from oracle import orc
    sample = bodies[:min(len(bodies), 16384)]
    arena = np.frombuffer(b"".join(sample), dtype=np.uint8)
    lens = np.fromiter((len(b) for b in sample), dtype=np.uint32, count=len(sample))
    offs = np.zeros(len(sample), dtype=np.uint64)
    offs[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
    modes = np.full(len(sample), mode, dtype=np.uint8)
    orc.bench_run(arena, offs, lens, modes, threads)           # warm-up (page in, allocator)
    total_s, frames, passes = 0.0, 0, 0
    while total_s < target_s and passes < 64:
        secs, _, fr, _ = orc.bench_run(arena, offs, lens, modes, threads)
        total_s += secs; frames += fr; passes += 1
    return {"value": frames / total_s, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{len(sample)} streams of the workload x {passes} passes, {arena.size / 1e6:.0f} MB per pass, "
                      f"{total_s:.1f} s of CPU wall time; oracle/sse_oracle.c (C restatement; the reference is Go and no Go "
                      f"toolchain exists on this box)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) with all host threads, rank 0 only."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    bodies, mode, _ = build_workload(min(args.streams, 16384), 0, args.workload)
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(bodies, mode, threads, target_s=0.5)
    for _ in range(args.steps):
        vals.append(cpu_baseline(bodies, mode, threads, target_s=2.0))
    v = statistics.mean(x["value"] for x in vals)
    cb = dict(vals[-1]); cb["value"] = v
    chunks_per_step = None
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: mixed cohere/groq/anthropic/ollama SSE streams, mode R (reframe + side-band), "
                                   f"bounded sample of {len(bodies)} streams per step", "threads": threads},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=65536, help="concurrent streams per GPU")
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--mode", type=int, default=None, help="override the workload's mode bits (0 P, 2 P+parse, 3 R+parse)")
    ap.add_argument("--flags", type=int, default=0, help="sse_config.flags (1 v1 kernel, 2 fused v2 kernel, 8 copy-out)")
    ap.add_argument("--n-content", type=int, default=None, help="content deltas per stream (default: the config's 7)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL printf()s its version banner to stdout when the communicator is created
        # (NCCL_DEBUG=VERSION/INFO), so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from inference_gateway_b200 import SseEngine, shard as sh
    bodies, mode, n_events = build_workload(args.streams, rank, args.workload, args.n_content)
    if args.mode is not None:
        mode = args.mode
    in_payload = sum(map(len, bodies))
    eng = SseEngine(device=local_rank, max_conns=len(bodies), bytes_per_batch=in_payload, n_slots=N_SLOTS, carry_slot_bytes=16384, flags=args.flags)
    # a non-default torch stream: its handle is passed to the library so that the kernels, the resets and the
    # CUDA events of the timed region all live on the same stream (handle 0 would mean "library stream")
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    slot, arena, segs = eng.acquire()
    n_segs, in_bytes, _ = fill_slot(eng, arena, segs, bodies, mode)
    eng.upload(slot, n_segs, in_bytes, stream)
    torch.cuda.synchronize()

    def step():
        eng.reset_all(stream)          # a step = a fresh population of connections
        eng.launch(slot, n_segs, stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    res = eng.download(slot, stream)   # untimed: counts for the report + sanity
    counts = dict(frames=int(res.raw.n_frames), recs=int(res.raw.n_recs), tcs=int(res.raw.n_tcs), usages=int(res.raw.n_usages),
                  out_bytes=int(res.raw.out_bytes), text_bytes=int(res.raw.text_bytes), runs=int(res.raw.n_runs),
                  in_bytes=in_payload, segs=n_segs, events=n_events)
    valid_frames = int(res.segs["frame_count"].astype(np.int64).sum()) + sum(
        int(res.runs["frame_count"][j]) for j in range(int(res.raw.n_runs)))
    counts["frames"] = valid_frames        # frames of lines after a terminating chunk are allocated but cut from the result
    # bytes of the delivered frames (b_out of SURVEY 8(d)), wherever they live: materialised in the out arena or spans of the input
    cs = np.concatenate([[0], np.cumsum(res.frames["len"].astype(np.int64))])
    ff, fc = res.segs["frame_first"].astype(np.int64), res.segs["frame_count"].astype(np.int64)
    frame_bytes = int((cs[ff + fc] - cs[ff]).sum())
    for j in range(int(res.raw.n_runs)):
        a, b = int(res.runs["frame_first"][j]), int(res.runs["frame_count"][j])
        frame_bytes += int(cs[a + b] - cs[a])
    counts["frame_bytes"] = frame_bytes
    counts["zero_copy_frames"] = int(np.count_nonzero(res.frames["off"] >= res.in_base))
    terminated = int(np.count_nonzero(res.segs["flags"] & 1))
    ok_recs = int(np.count_nonzero(res.recs["flags"] & 1))

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    # nvidia-smi needs a few hundred ms before its first sample: keep the GPU under the same load (untimed) meanwhile,
    # so that the clock samples bracket the timed region and are all taken under load
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 1.2:
        step()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()

    # ---- end to end through the public C ABI: pinned H2D + kernel + D2H every step, two slots in flight
    e2e = None
    if not args.no_e2e:
        eng.release(slot)
        slots = []
        for _ in range(N_SLOTS):
            s2, a2, g2 = eng.acquire()
            fill_slot(eng, a2, g2, bodies, mode)
            slots.append(s2)
        for s2 in slots:
            eng.release(s2)
        h2d = in_bytes + n_segs * 16
        d2h = (counts["out_bytes"] + 8 * counts["frames"] + 32 * counts["recs"] + 48 * counts["tcs"] + 24 * counts["usages"]
               + counts["text_bytes"] + 20 * counts["runs"] + 32 * n_segs + 64)

        def e2e_steps(k):
            inflight = []
            frames = 0
            for i in range(k):
                s2, _, _ = eng.acquire()      # pinned staging already holds the step's input
                eng.reset_all(0)
                eng.submit(s2, n_segs, in_bytes)
                inflight.append(s2)
                if len(inflight) == N_SLOTS:
                    s_old = inflight.pop(0)
                    r = eng.collect(s_old)
                    frames += int(r.raw.n_frames)
                    eng.release(s_old)
            for s_old in inflight:
                r = eng.collect(s_old)
                frames += int(r.raw.n_frames)
                eng.release(s_old)
            return frames

        e2e_steps(N_SLOTS)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        k_e2e = max(4, min(args.steps, 12))
        t0 = time.perf_counter()
        fr = e2e_steps(k_e2e)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_s = sh.max_over_ranks(e2e_s, world)
        fr_all = sh.reduce_counters({"f": fr}, world)["f"]
        e2e = {"value": fr_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": k_e2e, "ms_per_step": 1e3 * e2e_s / k_e2e,
               "path": f"sse_acquire/sse_submit/sse_collect/sse_release, {N_SLOTS} slots in flight, host buffers pinned"}

    dev_ms = sh.max_over_ranks(dev_ms, world)
    tot = sh.reduce_counters(counts, world)
    chunks_per_step = tot["frames"]
    value = chunks_per_step * args.steps / (dev_ms / 1e3)

    # ---- roofline of the stream kernel (rank 0's launch): algorithmic bytes / measured duration
    # b_in read once + b_out delivered once + frame table + records (+ per-segment descriptors and state).
    # moved_bytes is what the zero-copy path really has to touch: frames that are spans of the input are never written.
    side = (8 * counts["frames"] + 32 * counts["recs"] + 48 * counts["tcs"] + 24 * counts["usages"] + counts["text_bytes"]
            + (16 + 32 + 16) * n_segs)
    alg_bytes = counts["in_bytes"] + counts["frame_bytes"] + side
    moved_bytes = counts["in_bytes"] + counts["out_bytes"] + side
    kern_s = (dev_ms / 1e3) / args.steps
    peak, peak_src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp["hbm_gbs"]), "measured"
    except (OSError, KeyError, ValueError):
        pass
    achieved = alg_bytes / kern_s / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("workload") == args.workload and tj.get("streams") == args.streams and args.mode == 3 and not args.flags:
            traffic = tj.get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "split pipeline: sse_stream_kernel<produce> + bucket sort (3 small kernels) + sse_decode_kernel + sse_finalize_kernel (whole step)",
                "alg_bytes_per_launch": alg_bytes, "moved_bytes_per_launch": moved_bytes,
                "zero_copy_frames": counts["zero_copy_frames"], "kernel_ms": 1e3 * kern_s,
                "hbm_read_frac": counts["in_bytes"] / kern_s / 1e9 / peak}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {args.streams} concurrent mixed cohere/groq/anthropic/ollama streams per GPU incl. "
                               "tool_calls deltas, mode R (MCP reframe + JSON side-band), every stream complete in one micro-batch",
                   "streams_per_gpu": args.streams, "total_streams": args.streams * world,
                   "sse_events_per_step": tot["events"], "chunks_emitted_per_step": tot["frames"],
                   "records_decoded_per_step": tot["recs"], "mean_event_bytes": tot["in_bytes"] / max(1, tot["events"]),
                   "input_bytes_per_step": tot["in_bytes"], "sharding": "hash(conn_id) % n_gpus, no collective",
                   "l2": "working set (input + result tables > 350 MB per GPU) exceeds the 126 MB L2; no explicit flush",
                   "streams_terminated": terminated, "records_json_ok": ok_recs},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if rank == 0 and not args.no_cpu_baseline and world >= 1:
        try:
            line["cpu_baseline"] = cpu_baseline(bodies, mode, os.cpu_count() or 1)
        except Exception as ex:  # the checker library is test infrastructure; report, do not hide
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
