"""Shared parity harness: drives streams through the GPU path in randomly cut micro-batches and compares
every frame and every side-band record with the CPU oracle (oracle/ is the checker, never the product)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from inference_gateway_b200 import _abi as A
from oracle import orc


class ConnOut:
    def __init__(self):
        self.frames = []       # emitted frames (bytes) in order
        self.recs = []         # decoded records (dicts) in order
        self.terminated = False
        self.carry_len = 0
        self.flags = 0
        self.agent = None
        self.tele = None


def rec_to_dict(res, ri: int) -> dict:
    r = res.recs[ri]
    fl = int(r["flags"])
    d = dict(flags=fl, json_ok=bool(fl & A.F_JSON_OK), done=bool(fl & A.F_DONE_LINE),
             done_exact=bool(fl & A.F_DONE_EXACT), terminates=bool(fl & A.F_TERMINATES),
             depth_limit=bool(fl & A.F_DEPTH_LIMIT), too_long=bool(fl & A.F_TOO_LONG), frame=int(r["frame"]), payload_len=int(r["payload_len"]))
    if d["json_ok"]:
        d["n_choices"] = int(r["n_choices"])
        d["finish"] = (fl & A.F_FINISH_MASK) >> A.F_FINISH_SHIFT
        d["usage"] = None
        if fl & A.F_HAS_USAGE:
            u = res.usages[int(r["usage"])]
            d["usage"] = (int(u["prompt"]), int(u["completion"]), int(u["total"]))
        d["content"] = res.span(int(r["content_off"]), int(r["content_len"]), bool(fl & A.F_CONTENT_TEXT))
        d["tc_nonnil"] = bool(fl & A.F_TC_NONNIL)
        d["tc_valid"] = bool(fl & A.F_TC_VALID)
        tcs = []
        t = int(r["tc_first"])
        for _ in range(int(r["tc_count"])):
            assert t != A.NONE, "tool-call chain shorter than tc_count"
            tc = res.tcs[t]
            tf = int(tc["flags"])
            tcs.append(dict(
                index=int(tc["index"]),
                id=res.span(int(tc["id_off"]), int(tc["id_len"]), bool(tf & A.TC_ID_TEXT)) if tf & A.TC_HAS_ID else None,
                type=res.span(int(tc["type_off"]), int(tc["type_len"]), bool(tf & A.TC_TYPE_TEXT)) if tf & A.TC_HAS_TYPE else None,
                function=bool(tf & A.TC_HAS_FUNC),
                name=res.span(int(tc["name_off"]), int(tc["name_len"]), bool(tf & A.TC_NAME_TEXT)),
                args=res.span(int(tc["args_off"]), int(tc["args_len"]), bool(tf & A.TC_ARGS_TEXT))))
            t = int(tc["next"])
        d["tcs"] = tcs
    return d


def chunk_to_dict(ck: orc.ChunkView) -> dict:
    d = dict(json_ok=ck.json_ok)
    if ck.json_ok:
        d.update(n_choices=min(ck.n_choices, 0xFFFF), finish=ck.finish, usage=ck.usage, content=ck.content,
                 tc_nonnil=ck.tool_calls_nonnil, tc_valid=ck.has_valid_tool_call,
                 tcs=[dict(index=t.index, id=t.id, type=t.type, function=t.function, name=t.name, args=t.args)
                      for t in ck.tool_calls])
    return d


def run_streams(engine, streams, modes, n_batches=1, seed=0, with_folds=False):
    """streams: list of bytes (one upstream body per connection); modes: list of mode bits.
    Feeds every stream in n_batches randomly cut pieces (simulated TCP segmentation), one micro-batch per
    piece index. Returns list[ConnOut]."""
    from inference_gateway_b200.synth import random_cuts
    rng = np.random.default_rng(seed)
    L = engine.L
    n = len(streams)
    outs = [ConnOut() for _ in range(n)]
    engine.reset_all()
    pieces = [random_cuts(rng, s, n_batches) for s in streams]
    if with_folds:
        for o in outs:
            o.agent = L.sse_agent_new()
            o.tele = L.sse_telemetry_new()
    rounds = max(len(p) for p in pieces) if pieces else 0
    for b in range(rounds):
        items, who = [], []
        for c in range(n):
            if b < len(pieces[c]):
                items.append((c, modes[c], pieces[c][b]))
                who.append(c)
        if not items:
            continue
        slot, res = engine.process(items)
        try:
            for i, c in enumerate(who):
                o = outs[c]
                o.frames.extend(res.seg_frames(i))
                o.recs.extend(rec_to_dict(res, ri) for ri in res.seg_recs(i))
                sr = res.segs[i]
                o.carry_len = int(sr["carry_len"])
                o.flags |= int(sr["flags"])
                if int(sr["flags"]) & A.SEG_TERMINATED:
                    o.terminated = True
                if with_folds:
                    A.check(L.sse_agent_feed(o.agent, C.byref(res.raw), i), "sse_agent_feed")
                    A.check(L.sse_telemetry_feed(o.tele, C.byref(res.raw), i), "sse_telemetry_feed")
        finally:
            engine.release(slot)
    return outs


def _bytes(b: A.Bytes) -> bytes:
    return C.string_at(b.p, b.n) if b.n else b""


def agent_results(L, fold):
    content = _bytes(L.sse_agent_content(fold))
    has = bool(L.sse_agent_has_tool_calls(fold))
    fin = C.c_int()
    term = bool(L.sse_agent_terminated(fold, C.byref(fin)))
    arr = (A.ToolCall * 64)()
    n = L.sse_agent_tool_calls(fold, arr, 64)
    calls = [dict(id=_bytes(arr[i].id), type=_bytes(arr[i].type), name=_bytes(arr[i].name), args=_bytes(arr[i].arguments))
             for i in range(min(n, 64))]
    return content, has, term, fin.value, calls


def telemetry_results(L, fold):
    u = A.Usage()
    arr = (A.ToolCall * 64)()
    n = C.c_size_t()
    rc = L.sse_telemetry_finish(fold, C.byref(u), arr, 64, C.byref(n))
    calls = [dict(id=_bytes(arr[i].id), type=_bytes(arr[i].type), name=_bytes(arr[i].name), args=_bytes(arr[i].arguments))
             for i in range(min(n.value, 64))]
    return rc, (u.prompt_tokens, u.completion_tokens, u.total_tokens), calls


def check_stream(body: bytes, mode: int, o: ConnOut, label=""):
    """Bit-exact comparison of one connection's GPU output with the oracle."""
    if mode & A.MODE_R:
        v = orc.reframe(body)
        exp_frames = [l.out for l in v.lines if l.kind == orc.L_EMITTED]
        exp_recs = []
        for l in v.lines:
            if l.kind == orc.L_EMITTED or l.kind == orc.L_DONE:
                d = chunk_to_dict(l.chunk)
                d["done"] = l.kind == orc.L_DONE
                d["done_exact"] = False
                exp_recs.append(d)
            elif l.kind == orc.L_DONE_EXACT:
                exp_recs.append(dict(json_ok=False, done=True, done_exact=True))
        assert o.terminated == v.terminated, f"{label}: terminated {o.terminated} != {v.terminated}"
    else:
        v = orc.passthrough(body, parse=bool(mode & A.MODE_PARSE))
        exp_frames = [l.out for l in v.lines]
        exp_recs = []
        for l in v.lines:
            if l.chunk is not None:
                d = chunk_to_dict(l.chunk)
                d["done"] = False
                d["done_exact"] = False
                exp_recs.append(d)
        assert b"".join(o.frames) == v.out, f"{label}: passthrough bytes differ"
    assert len(o.frames) == len(exp_frames), f"{label}: {len(o.frames)} frames, oracle {len(exp_frames)}"
    for i, (g, e) in enumerate(zip(o.frames, exp_frames)):
        assert g == e, f"{label}: frame {i} differs:\n gpu={g[:120]!r}\n ref={e[:120]!r}"
    assert len(o.recs) == len(exp_recs), f"{label}: {len(o.recs)} recs, oracle {len(exp_recs)}"
    for i, (g, e) in enumerate(zip(o.recs, exp_recs)):
        assert not g["too_long"], f"{label}: rec {i} was not decoded (SSE_F_TOO_LONG)"
        if g["depth_limit"]:      # nesting deeper than 128 levels: the documented limit, reported as not JSON_OK
            assert not g["json_ok"]
            continue
        for k, ev in e.items():
            assert g[k] == ev, f"{label}: rec {i} field {k}: gpu={g[k]!r} ref={ev!r}"
    if not o.terminated and not (o.flags & A.SEG_DEAD):
        assert o.carry_len == v.tail_len, f"{label}: carry {o.carry_len} != tail {v.tail_len}"
    return v
