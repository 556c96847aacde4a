"""CPU: the host-side folds (sse_agent_*, sse_telemetry_*) evaluated on records synthesised from the oracle's
per-line results must reproduce the oracle's restatement of agent.go:377-481 and telemetry.go:190-277."""
import ctypes as C

import numpy as np
import pytest

from inference_gateway_b200 import _abi as A
from inference_gateway_b200 import synth
from oracle import orc
from tests.util import agent_results, telemetry_results


class FakeResult:
    """Builds an sse_result (one segment) from oracle line views, the way the kernel lays it out."""

    def __init__(self, view: orc.StreamView, mode_r: bool, in_arena: bool = False):
        # in_arena: every frame is a span of the input arena (zero-copy frames, offsets >= in_base), none is in out
        IN_BASE = 4096 if in_arena else 0xFFFFFFFF
        out, text = bytearray(), bytearray(b"\0")
        frames, recs, tcs, usages = [], [], [], []

        def put(b):
            o = len(text); text.extend(b); return o, len(b)

        for l in view.lines:
            emitted = l.kind == orc.L_EMITTED
            if l.kind == orc.L_UNREAD:
                break
            fidx = A.NONE
            if emitted:
                fidx = len(frames); frames.append(((IN_BASE if in_arena else 0) + len(out), len(l.out))); out.extend(l.out)
            ck = l.chunk
            if l.kind == orc.L_DONE_EXACT:
                recs.append((A.NONE, A.F_DONE_LINE | A.F_DONE_EXACT, 0, 0, A.NONE, 0, 0, A.NONE, 6))
                continue
            if ck is None:
                continue
            fl = A.F_DONE_LINE if l.kind == orc.L_DONE else 0
            co = cl = 0; tcf = A.NONE; tcc = 0; nch = 0; ui = A.NONE
            if ck.json_ok:
                fl |= A.F_JSON_OK | (ck.finish << A.F_FINISH_SHIFT)
                nch = min(ck.n_choices, 0xFFFF)
                if ck.usage is not None:
                    fl |= A.F_HAS_USAGE; ui = len(usages); usages.append(ck.usage)
                if ck.content:
                    co, cl = put(ck.content); fl |= A.F_CONTENT_TEXT
                if ck.tool_calls_nonnil:
                    fl |= A.F_TC_NONNIL
                if ck.has_valid_tool_call:
                    fl |= A.F_TC_VALID
                if mode_r and emitted and nch > 0 and ck.finish in (orc.FIN_STOP, orc.FIN_TOOL_CALLS):
                    fl |= A.F_TERMINATES
                tcc = len(ck.tool_calls)
                for k, t in enumerate(ck.tool_calls):
                    tf = A.TC_ID_TEXT | A.TC_TYPE_TEXT | A.TC_NAME_TEXT | A.TC_ARGS_TEXT
                    tf |= (A.TC_HAS_ID if t.id is not None else 0) | (A.TC_HAS_TYPE if t.type is not None else 0)
                    tf |= A.TC_HAS_FUNC if t.function else 0
                    io, il = put(t.id or b""); to, tl = put(t.type or b""); no, nl = put(t.name); ao, al = put(t.args)
                    idx = len(tcs)
                    if k == 0:
                        tcf = idx
                    else:
                        tcs[-1][2] = idx
                    tcs.append([t.index, tf, A.NONE, io, il, to, tl, no, nl, ao, al])
            recs.append((fidx, fl, co, cl, tcf, tcc, nch, ui, 0))
        self.keep = []

        def arr(cls, rows):
            a = (cls * max(1, len(rows)))()
            for i, r in enumerate(rows):
                a[i] = cls(*r)
            self.keep.append(a)
            return a

        res = A.Result()
        res.status = 0
        res.n_segs = 1
        res.n_frames, res.n_recs, res.n_tcs, res.n_usages = len(frames), len(recs), len(tcs), len(usages)
        res.out_bytes, res.text_bytes = len(out), len(text)
        ob = (C.c_uint8 * max(1, len(out))).from_buffer_copy(bytes(out) or b"\0")
        tb = (C.c_uint8 * len(text)).from_buffer_copy(bytes(text))
        self.keep += [ob, tb]
        res.text = C.cast(tb, C.POINTER(C.c_uint8))
        res.in_base = IN_BASE
        if in_arena:
            res.in_ = C.cast(ob, C.POINTER(C.c_uint8)); res.out_bytes = 0
        else:
            res.out = C.cast(ob, C.POINTER(C.c_uint8))
        res.frames = C.cast(arr(A.Frame, frames), C.POINTER(A.Frame))
        res.recs = C.cast(arr(A.Rec, recs), C.POINTER(A.Rec))
        res.tcs = C.cast(arr(A.Tc, [tuple(t) for t in tcs]), C.POINTER(A.Tc))
        res.usages = C.cast(arr(A.Usage, usages), C.POINTER(A.Usage))
        res.runs = C.cast(arr(A.Run, []), C.POINTER(A.Run))
        seg = A.SegResult()
        seg.run = A.Run(0, len(frames), 0, len(recs), A.NONE)
        sa = (A.SegResult * 1)(seg)
        self.keep.append(sa)
        res.segs = C.cast(sa, C.POINTER(A.SegResult))
        self.res = res


def _streams():
    streams, _ = synth.make_config("C4", n_streams=120)
    extra = [
        b'data: {"choices":[{"delta":{"tool_calls":[{"index":0,"function":{"name":"a","arguments":"x"}},{"index":0,"function":{"arguments":"y"}}]}}]}\n\n'
        b'data: {"choices":[{"delta":{"tool_calls":[{"index":2,"id":"late","function":{"name":"gap"}}]}}]}\n\n'
        b'data: {"choices":[{"delta":{"content":"has [DONE] inside","tool_calls":[{"index":1,"id":"swallowed","function":{"name":"s"}}]}}]}\n\n'
        b'data: {"choices":[{"delta":{},"finish_reason":"tool_calls"}]}\n\ndata: [DONE]\n\n',
        b'data: {"choices":[{"delta":{"tool_calls":[{"index":0,"id":"c","function":{"name":"n","arguments":"{}"}}]}}]}\n\ndata: [DONE]\n\n'
        b'data: {"choices":[{"delta":{"tool_calls":[{"index":1,"id":"after done","function":{"name":"ignored"}}]}}]}\n\n',
        # a bare "[DONE]" line (no "data: " prefix) is swallowed (agent.go:181-184) but parseStreamingToolCalls only breaks on
        # the exact line "data: [DONE]" (agent.go:394-396): the tool call that follows still counts
        b'data: {"choices":[{"delta":{"tool_calls":[{"index":0,"id":"a","function":{"name":"first","arguments":"{"}}]}}]}\n\n[DONE]\n\n'
        b'data: {"choices":[{"delta":{"tool_calls":[{"index":0,"function":{"arguments":"}"}},{"index":1,"id":"b","function":{"name":"second"}}]}}]}\n\n'
        b'data: {"choices":[{"delta":{},"finish_reason":"tool_calls"}]}\n\n',
        b'  [DONE]  \n\ndata: {"choices":[{"delta":{"tool_calls":[{"index":0,"id":"x","function":{"name":"n"}}]}}]}\n\n',
    ]
    return [b for b, _, _ in streams] + extra


@pytest.mark.parametrize("in_arena", [False, True])
def test_agent_fold_matches_oracle(in_arena):
    L = A.load()
    for body in _streams():
        v = orc.reframe(body)
        fr = FakeResult(v, True, in_arena)
        f = L.sse_agent_new()
        A.check(L.sse_agent_feed(f, C.byref(fr.res), 0), "feed")
        content, has, term, fin, calls = agent_results(L, f)
        L.sse_agent_free(f)
        assert content == v.acc_content and has == v.has_tool_calls and term == v.terminated
        assert calls == orc.parse_tool_calls(v.builder), body[:80]


@pytest.mark.parametrize("in_arena", [False, True])
def test_telemetry_fold_matches_oracle(in_arena):
    L = A.load()
    bodies = _streams() + [b"data: {\"usage\":{\"prompt_tokens\":1,\"completion_tokens\":2,\"total_tokens\":3}}\n\n" + b"data: {\"choices\":[]}\n\n" * k + b"data: [DONE]\n\n"
                           for k in range(5)]
    bodies += [b"\n\ndata: {\"usage\":{\"prompt_tokens\":4}}\n\n\ndata: {\"usage\":{\"prompt_tokens\":9}}\n\n", b"", b"\n", b"\n\n\n\n",
               b"event: x\ndata: {\"usage\":{\"prompt_tokens\":4}}\n\n", b"data: {\"usage\":{\"prompt_tokens\":5}}\n"]
    for body in bodies:
        v = orc.passthrough(body, parse=True)
        fr = FakeResult(v, False, in_arena)
        f = L.sse_telemetry_new()
        A.check(L.sse_telemetry_feed(f, C.byref(fr.res), 0), "feed")
        rc, usage, calls = telemetry_results(L, f)
        L.sse_telemetry_free(f)
        eusage, ecalls = orc.telemetry(v.out)
        assert (usage, calls) == (eusage, ecalls), body[:100]


def test_mcp_path_telemetry_counts_the_final_done_frame():
    """On the MCP path the client body is the forwarded frames plus the agent's own "data: [DONE]\\n\\n" (agent.go:140-143);
    telemetry.go:195-198 looks at the last 4 pieces of THAT body, so the host feeds the synthetic frame too."""
    L = A.load()
    done = b"data: [DONE]\n\n"
    usage_ev = b'data: {"choices":[],"usage":{"prompt_tokens":7,"completion_tokens":8,"total_tokens":15}}\n\n'
    filler = b'data: {"choices":[{"delta":{"content":"x"}}]}\n\n'
    stop = b'data: {"choices":[{"delta":{},"finish_reason":"stop"}]}\n\n'
    seen = set()
    for k in range(4):          # usage event followed by k more events and the terminating one
        body = usage_ev + filler * k + stop
        v = orc.reframe(body)
        fr = FakeResult(v, True)
        f = L.sse_telemetry_new()
        A.check(L.sse_telemetry_feed(f, C.byref(fr.res), 0), "feed")
        A.check(L.sse_telemetry_feed_bytes(f, done, len(done)), "feed_bytes")
        rc, usage, calls = telemetry_results(L, f)
        L.sse_telemetry_free(f)
        eusage, ecalls = orc.telemetry(v.out + done)
        assert rc == 0 and (usage, calls) == (eusage, ecalls), k
        seen.add(usage)
    assert seen == {(7, 8, 15), (0, 0, 0)}         # the window really moves: with enough later pieces the usage falls out
    f = L.sse_telemetry_new()
    assert L.sse_telemetry_feed_bytes(f, b"no newline", 10) == A.SSE_ERR_ARG
    L.sse_telemetry_free(f)


def test_folds_report_undecoded_records():
    """A record flagged TOO_LONG / DEPTH_LIMIT is a line the reference would have decoded: the folds say so instead of
    skipping it silently."""
    L = A.load()
    v = orc.reframe(b'data: {"choices":[{"delta":{"content":"a"}}]}\n\n')
    for flag in (A.F_TOO_LONG, A.F_DEPTH_LIMIT):
        fr = FakeResult(v, True)
        fr.res.recs[0].flags = flag
        f = L.sse_agent_new()
        assert L.sse_agent_feed(f, C.byref(fr.res), 0) == A.SSE_ERR_UNDECODED
        L.sse_agent_free(f)
        t = L.sse_telemetry_new()
        assert L.sse_telemetry_feed(t, C.byref(fr.res), 0) == A.SSE_ERR_UNDECODED
        L.sse_telemetry_free(t)
    assert b"not decoded" in L.sse_strerror(A.SSE_ERR_UNDECODED)
