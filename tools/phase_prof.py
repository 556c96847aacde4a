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
out = (C.c_ulonglong * 16)()
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
