"""Thin Python driver over the C ABI (one context = one GPU).

Plumbing only: it fills the pinned staging buffers the library lends out, submits, collects and turns the
result arrays into numpy views. All work happens in libssegpu.so; nothing here parses or splits bytes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi as A


def _np_view(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype, count=count)


FRAME_DT = np.dtype([("off", "<u4"), ("len", "<u4")])
REC_DT = np.dtype([("frame", "<u4"), ("flags", "<u4"), ("content_off", "<u4"), ("content_len", "<u4"),
                   ("tc_first", "<u4"), ("tc_count", "<u2"), ("n_choices", "<u2"), ("usage", "<u4"),
                   ("payload_len", "<u4")])
TC_DT = np.dtype([("index", "<i8"), ("flags", "<u4"), ("next", "<u4"), ("id_off", "<u4"), ("id_len", "<u4"),
                  ("type_off", "<u4"), ("type_len", "<u4"), ("name_off", "<u4"), ("name_len", "<u4"),
                  ("args_off", "<u4"), ("args_len", "<u4")])
USAGE_DT = np.dtype([("prompt", "<i8"), ("completion", "<i8"), ("total", "<i8")])
RUN_DT = np.dtype([("frame_first", "<u4"), ("frame_count", "<u4"), ("rec_first", "<u4"), ("rec_count", "<u4"),
                   ("next", "<u4")])
SEGRES_DT = np.dtype([("frame_first", "<u4"), ("frame_count", "<u4"), ("rec_first", "<u4"), ("rec_count", "<u4"),
                      ("next", "<u4"), ("carry_len", "<u4"), ("flags", "<u4"), ("reserved", "<u4")])
SEG_DT = np.dtype([("conn", "<u4"), ("in_off", "<u4"), ("in_len", "<u4"), ("mode", "u1"), ("provider", "u1"),
                   ("reserved", "<u2")])
assert FRAME_DT.itemsize == 8 and REC_DT.itemsize == 32 and TC_DT.itemsize == 48 and SEGRES_DT.itemsize == 32
assert SEG_DT.itemsize == 16 and RUN_DT.itemsize == 20


@dataclass
class BatchResult:
    """Views into the library's pinned result buffers (valid until the slot is released)."""
    raw: A.Result
    out: np.ndarray
    frames: np.ndarray
    recs: np.ndarray
    tcs: np.ndarray
    usages: np.ndarray
    text: np.ndarray
    runs: np.ndarray
    segs: np.ndarray
    in_arena: np.ndarray = None     # the batch's input arena; arena offsets >= in_base index it (zero-copy frames)
    in_base: int = 0xFFFFFFFF

    def at(self, off: int, length: int) -> bytes:
        """sse_at(): resolve an arena offset."""
        if off >= self.in_base:
            o = off - self.in_base
            return self.in_arena[o:o + length].tobytes()
        return self.out[off:off + length].tobytes()

    def seg_runs(self, i: int):
        s = self.segs[i]
        run = (int(s["frame_first"]), int(s["frame_count"]), int(s["rec_first"]), int(s["rec_count"]))
        nxt = int(s["next"])
        yield run
        while nxt != A.NONE:
            r = self.runs[nxt]
            yield (int(r["frame_first"]), int(r["frame_count"]), int(r["rec_first"]), int(r["rec_count"]))
            nxt = int(r["next"])

    def seg_frames(self, i: int) -> list:
        out = []
        for ff, fc, _, _ in self.seg_runs(i):
            for k in range(ff, ff + fc):
                f = self.frames[k]
                out.append(self.at(int(f["off"]), int(f["len"])))
        return out

    def seg_recs(self, i: int) -> list:
        out = []
        for _, _, rf, rc in self.seg_runs(i):
            out.extend(range(rf, rf + rc))
        return out

    def span(self, off: int, length: int, in_text: bool) -> bytes:
        return self.text[off:off + length].tobytes() if in_text else self.at(off, length)


class SseEngine:
    def __init__(self, device: int = 0, max_conns: int = 1024, bytes_per_batch: int = 1 << 20, **overrides):
        self.L = A.load()
        cfg = A.Config()
        self.L.sse_default_config(C.byref(cfg), max_conns, bytes_per_batch)
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self._ctx = C.c_void_p()
        A.check(self.L.sse_init(device, C.byref(cfg), C.byref(self._ctx)), "sse_init")
        self.device = device

    def close(self):
        if self._ctx:
            self.L.sse_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- slot level ------------------------------------------------------------------------------
    def acquire(self):
        slot = C.c_int()
        b = A.Batch()
        A.check(self.L.sse_acquire(self._ctx, C.byref(slot), C.byref(b)), "sse_acquire")
        arena = _np_view(b.in_arena, b.in_arena_bytes, np.uint8)
        segs = _np_view(C.cast(b.segs, C.POINTER(C.c_uint8)), b.max_segs * 16, np.uint8).view(SEG_DT)
        return slot.value, arena, segs

    def fill(self, arena, segs, items):
        """items: iterable of (conn, mode, bytes). Returns (n_segs, in_bytes)."""
        off = 0
        n = 0
        for conn, mode, data in items:
            ln = len(data)
            if ln:
                arena[off:off + ln] = np.frombuffer(data, dtype=np.uint8)
            segs[n] = (conn, off, ln, mode, 0, 0)
            off = (off + ln + 15) & ~15
            n += 1
        return n, off

    def submit(self, slot, n_segs, in_bytes):
        A.check(self.L.sse_submit(self._ctx, slot, n_segs, in_bytes), "sse_submit")

    def _wrap(self, res: A.Result) -> BatchResult:
        return BatchResult(
            res,
            _np_view(res.out, res.out_bytes, np.uint8),
            _np_view(C.cast(res.frames, C.POINTER(C.c_uint8)), res.n_frames * 8, np.uint8).view(FRAME_DT),
            _np_view(C.cast(res.recs, C.POINTER(C.c_uint8)), res.n_recs * 32, np.uint8).view(REC_DT),
            _np_view(C.cast(res.tcs, C.POINTER(C.c_uint8)), res.n_tcs * 48, np.uint8).view(TC_DT),
            _np_view(C.cast(res.usages, C.POINTER(C.c_uint8)), res.n_usages * 24, np.uint8).view(USAGE_DT),
            _np_view(res.text, res.text_bytes, np.uint8),
            _np_view(C.cast(res.runs, C.POINTER(C.c_uint8)), res.n_runs * 20, np.uint8).view(RUN_DT),
            _np_view(C.cast(res.segs, C.POINTER(C.c_uint8)), res.n_segs * 32, np.uint8).view(SEGRES_DT),
            _np_view(res.in_, self.cfg.in_arena_bytes, np.uint8), int(res.in_base),
        )

    def collect(self, slot) -> BatchResult:
        res = A.Result()
        rc = self.L.sse_collect(self._ctx, slot, C.byref(res))
        if rc == A.SSE_ERR_OVERFLOW:
            raise A.SseError(rc, f"sse_collect (overflow mask 0x{res.overflow:x}: 1 out/frames/recs, 2 tcs, 4 usages, 8 text, 16 runs)")
        A.check(rc, "sse_collect")
        return self._wrap(res)

    def release(self, slot):
        A.check(self.L.sse_release(self._ctx, slot), "sse_release")

    def reset_conn(self, conn: int):
        A.check(self.L.sse_reset_conn(self._ctx, conn), "sse_reset_conn")

    def reset_all(self, stream: int = 0):
        A.check(self.L.sse_reset_all(self._ctx, C.c_void_p(stream)), "sse_reset_all")

    # -- device-resident stages (bench) ------------------------------------------------------------
    def upload(self, slot, n_segs, in_bytes, stream: int = 0):
        A.check(self.L.sse_upload(self._ctx, slot, n_segs, in_bytes, C.c_void_p(stream)), "sse_upload")

    def launch(self, slot, n_segs, stream: int = 0):
        A.check(self.L.sse_launch(self._ctx, slot, n_segs, C.c_void_p(stream)), "sse_launch")

    def download(self, slot, stream: int = 0) -> BatchResult:
        res = A.Result()
        rc = self.L.sse_download(self._ctx, slot, C.byref(res), C.c_void_p(stream))
        if rc == A.SSE_ERR_OVERFLOW:
            raise A.SseError(rc, f"sse_download (overflow mask 0x{res.overflow:x}: 1 out/frames/recs, 2 tcs, 4 usages, 8 text, 16 runs)")
        A.check(rc, "sse_download")
        return self._wrap(res)

    def launch_count(self) -> int:
        n = C.c_uint64()
        A.check(self.L.sse_launch_count(self._ctx, C.byref(n)), "sse_launch_count")
        return n.value

    # -- convenience: one synchronous batch ---------------------------------------------------------
    def process(self, items):
        """items: list of (conn, mode, bytes). Returns a BatchResult; caller must release(result_slot)."""
        slot, arena, segs = self.acquire()
        n, nbytes = self.fill(arena, segs, items)
        self.submit(slot, n, nbytes)
        r = self.collect(slot)
        return slot, r
