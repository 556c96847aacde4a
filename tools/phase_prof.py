"""Per-phase cycle counters of the fused kernel (build with SSE_NVCC_DEFS=-DSSE_PROF). Sums over CTAs of thread 0's clock64 deltas."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inference_gateway_b200 import SseEngine, _abi as A, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
streams, _ = synth.make_config(wl, n_streams=n)
bodies = [b for b, _, _ in streams]
tot = sum(map(len, bodies))
eng = SseEngine(device=0, max_conns=n, bytes_per_batch=tot, n_slots=1, carry_slot_bytes=16384)
slot, arena, segs = eng.acquire()
ns, nb = eng.fill(arena, segs, [(i, A.MODE_R | A.MODE_PARSE, b) for i, b in enumerate(bodies)])
eng.upload(slot, ns, nb)
out = (C.c_ulonglong * 160)()
names = ["setup+load", "stage1a", "enum+classify", "alloc+sort", "barrier after stage2", "frames+serialize", "runs", "finish_segment", "stage2 (thread 0's warp)"]
import torch
for it in range(4):
    eng.reset_all(); eng.launch(slot, ns); torch.cuda.synchronize()
    eng.L.sse_prof_read(out)
v = list(out)
s = sum(v[:8])
print(f"{wl} {n} streams, {tot/1e6:.1f} MB")
for i, nm in enumerate(names):
    print(f"{nm:28s} {v[i]/1e6:10.1f} Mcycles {v[i]/s*100 if i < 8 else v[i]/s*100:5.1f}%")

print("per-warp stage 2: Mcycles / loop iterations (x4 steps) / cycles per iteration")
for w_ in range(16):
    c_, i_ = v[16 + w_], v[32 + w_]
    print(f"  warp {w_:2d}: {c_/1e6:8.1f} Mcyc {i_/1e6:8.2f} Miter {c_/max(i_,1):8.1f} cyc/iter")

print("templates: verify attempts %d, hits %d, automaton lines %d, templates built %d (%d words), recordings %d" % tuple(v[48:54]))

print("stage 2 sub-phases per warp (Mcycles): setup / find(raw key) / clean key + find / apply")
for w_ in range(16):
    print("  warp %2d: " % w_ + " ".join("%8.1f" % (v[64 + k * 16 + w_] / 1e6) for k in range(4)))
