"""Prints the skeleton templates the fused kernel has learnt on a workload (debugging aid)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from inference_gateway_b200 import SseEngine, _abi as A, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
streams, _ = synth.make_config(wl, n_streams=n)
bodies = [b for b, _, _ in streams]
eng = SseEngine(device=0, max_conns=n, bytes_per_batch=sum(map(len, bodies)), n_slots=1, carry_slot_bytes=16384)
slot, arena, segs = eng.acquire()
ns, nb = eng.fill(arena, segs, [(i, A.MODE_R | A.MODE_PARSE, b) for i, b in enumerate(bodies)])
eng.upload(slot, ns, nb)
for it in range(3):
    eng.reset_all(); eng.launch(slot, ns); torch.cuda.synchronize()
buf = (C.c_uint32 * (257 + 4096))()
got = eng.L.sse_debug_tcache(eng._ctx, buf, len(buf))
w = np.frombuffer(buf, dtype=np.uint32)
used = int(w[0]); heads = w[1:257]; st = w[257:]
print("words used", used)
for key in range(256):
    off = int(heads[key]); chain = []
    while off:
        chain.append(off); off = int(st[off]) & 0xFFFF
    for off in chain:
        T = st[off:]
        n_items = int(T[3]) >> 16; tc = int(T[1]) >> 24; litb = int(T[1]) & 0xFFFF
        items = T[6:6 + n_items]
        lw = T[6 + n_items + (4 if tc else 0):]
        raw = lw.tobytes()
        out = []; p = 0
        for it in items:
            lit = int(it) & 0xFFFF; kind = (int(it) >> 16) & 0xFF; op = int(it) >> 24
            out.append(raw[p:p + lit].decode("latin1")); p += (lit + 3) & ~3
            out.append({0: "", 1: "<S%d>" % op, 2: "<I%d>" % op}[kind])
        print(f"key {key:3d} off {off:4d} items {n_items:2d} tc {tc} lit {litb:3d}: {''.join(out)}")
