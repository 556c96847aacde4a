"""GPU: the reference's own streaming tests, re-stated against the host-side mirror of its interfaces
(include/sse_gateway.h). Each test cites the Go test it follows; inputs are the reference's fixtures."""
import json
import os

import pytest

from inference_gateway_b200 import _abi as A
from inference_gateway_b200.gateway import Gateway

pytestmark = pytest.mark.gpu
GOLD = {f["name"]: f for f in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.json")))["fixtures"]}
R = A.MODE_R | A.MODE_PARSE


@pytest.fixture(scope="module")
def gw():
    g = Gateway(device=0, max_conns=64, bytes_per_batch=1 << 20)
    yield g
    g.close()


def run_agent_iteration(gw, elements, cut=None):
    """mockProvider.StreamChatCompletions returns a channel fed with `elements`; the agent forwards to the middleware
    channel. Here the elements arrive as upstream bytes (each followed by '\\n'), optionally cut into odd pieces."""
    sid = gw.stream_chat_completions(R)
    body = ("\n".join(elements) + "\n").encode()
    pieces = [body] if not cut else [body[i:i + cut] for i in range(0, len(body), cut)]
    out = []

    def drain():
        try:
            while True:
                e = gw.agent_recv(sid)
                if e is None:
                    return
                out.append(e)
        except EOFError:      # middleware channel closed: the iteration is over (after the single [DONE])
            return

    for p in pieces:
        if gw.upstream_write(sid, p) == len(p):
            gw.pump()
        drain()
    gw.upstream_close(sid)
    gw.pump()
    drain()
    state = gw.agent_state(sid)
    gw.release(sid)
    return out, state


@pytest.mark.parametrize("cut", [None, 37])
def test_agent_run_with_stream_no_tool_calls(gw, cut):
    """tests/mcp_agent_test.go:522-598 "no tool calls streaming": chunks are forwarded, the concatenated delta.content is
    "Hello there!", the stream ends with one [DONE]."""
    fx = GOLD["agent_no_tool_calls"]
    out, (content, has_tc, term, fin, calls) = run_agent_iteration(gw, fx["iterations"][0], cut)
    assert content.decode() == "Hello there!"
    assert not has_tc and calls == [] and term and fin == 1
    assert out[-1] == b"data: [DONE]\n\n" and sum(e == b"data: [DONE]\n\n" for e in out) == 1
    assert len(out) == 5 and all(e.startswith(b"data: {") and e.endswith(b"\n\n") for e in out[:-1])
    joined = b"".join(json.loads(e[6:])["choices"][0]["delta"].get("content", "").encode() for e in out[:-1])
    assert joined == b"Hello there!"


@pytest.mark.parametrize("cut", [None, 101])
def test_agent_run_with_stream_two_iterations(gw, cut):
    """tests/mcp_agent_test.go:665-862: iteration 1 ends with finish_reason tool_calls and two parsed calls whose
    arguments are {"param":"value"} and {"action":"execute"} (:745-750); iteration 2 ends with stop."""
    fx = GOLD["agent_two_iterations_tool_calls"]
    out1, (c1, has1, term1, fin1, calls1) = run_agent_iteration(gw, fx["iterations"][0], cut)
    assert c1.decode() == "I'll use both tools to help you." and has1 and term1 and fin1 == 2
    assert [(c["id"], c["name"], json.loads(c["args"])) for c in calls1] == [
        (b"call_123", b"mcp_test_tool", {"param": "value"}), (b"call_456", b"mcp_other_tool", {"action": "execute"})]
    out2, (c2, has2, term2, fin2, calls2) = run_agent_iteration(gw, fx["iterations"][1], cut)
    assert c2.decode() == "Based on the tool results, both tools executed successfully!" and not has2 and fin2 == 1
    assert calls2 == []
    # the middleware writes every iteration's frames and the single terminal frame of the last one (mcp.go:253-299)
    assert out2[-1] == b"data: [DONE]\n\n"


def test_provider_channel_passthrough_and_backpressure(gw):
    """provider.go:307-340: one element per '\\n'-terminated line including the newline; the unterminated tail is dropped at
    EOF; a full channel (100 elements) stops accepting upstream bytes until the receiver drains it."""
    sid = gw.stream_chat_completions(A.MODE_P)
    assert gw.upstream_write(sid, b"data: a\n\ndata: b") == 16
    gw.pump()
    assert gw.recv(sid) == b"data: a\n" and gw.recv(sid) == b"\n" and gw.recv(sid) is None
    assert gw.upstream_write(sid, b"\n" * 150) == 150
    gw.pump()
    assert gw.upstream_write(sid, b"x\n") == 0            # 151 elements queued: back-pressure
    got = []
    while (e := gw.recv(sid)) is not None:
        got.append(e)
    assert got[0] == b"data: b\n" and len(got) == 150
    assert gw.upstream_write(sid, b"tail without newline") == 20
    gw.upstream_close(sid)
    gw.pump()
    with pytest.raises(EOFError):
        gw.recv(sid)
    gw.release(sid)


def test_mcp_writer_loop_terminal_frame_and_error_status(gw):
    """api/middlewares/mcp.go:253-299 over the agent's channel (tests/middlewares/mcp_test.go:768-918 pins one [DONE] at the
    end): every frame is written unchanged, the stream ends at the frame that is byte-equal to "data: [DONE]\\n\\n", and an
    upstream error object seen on the way sets 503 without touching the bytes."""
    sid = gw.stream_chat_completions(R)
    body = (b'data: {"choices":[{"index":0,"delta":{"content":"partial"},"finish_reason":null}]}\n\n'
            b'data: {"error": "upstream exploded"}\n\n'
            b'data: {"choices":[{"index":0,"delta":{},"finish_reason":"stop"}]}\n\n'
            b'data: {"choices":[],"usage":{"prompt_tokens":1,"completion_tokens":1,"total_tokens":2}}\n\ndata: [DONE]\n\n')
    assert gw.upstream_write(sid, body) == len(body)
    gw.pump()
    gw.upstream_close(sid)
    gw.pump()
    written, status_503, ended = gw.mcp_write_loop(sid)
    gw.release(sid)
    assert ended and status_503
    assert written == (b'data: {"choices":[{"index":0,"delta":{"content":"partial"},"finish_reason":null}]}\n\n'
                       b'data: {"error": "upstream exploded"}\n\n'
                       b'data: {"choices":[{"index":0,"delta":{},"finish_reason":"stop"}]}\n\n'
                       b'data: [DONE]\n\n')       # the usage-only chunk after the terminating one is never read (agent.go:235-242)


def test_raw_proxy_stream_loop(gw):
    """api/routes.go:129-232 handleStreamingRequest: every '\\n'-terminated line of the upstream body is written as is (SSE
    comments, event: lines, CRLF, blank separators, non-SSE bytes), in order; at EOF the unterminated tail is dropped
    (:187-195) and the loop ends."""
    sid = gw.proxy_stream()
    body = (b": keep-alive\n\nevent: message\r\ndata: {\"x\":1}\r\n\r\n" + b"data: " + b"z" * 5000 + b"\n\n"
            b"not sse at all\n\n\n" + b"data: [DONE]\n\nunterminated tail")
    written = bytearray()
    lines = []
    for i in range(0, len(body), 997):                    # upstream TCP segments
        piece = body[i:i + 997]
        assert gw.upstream_write(sid, piece) == len(piece)
        gw.pump()
        while (e := gw.proxy_step(sid)) is not None:
            lines.append(e); written += e
    gw.upstream_close(sid)
    gw.pump()
    with pytest.raises(EOFError):
        gw.proxy_step(sid)
    gw.release(sid)
    assert bytes(written) == body[:body.rfind(b"\n") + 1]
    assert lines == [l + b"\n" for l in body.split(b"\n")[:-1]]      # one write (+ flush) per ReadBytes line


def test_one_stream_of_blank_lines_cannot_starve_or_overflow_the_batch():
    """Result capacities are sized for the worst case of the bytes a batch can hold (sse_worst_case_config): a connection that
    sends nothing but '\\n' (one frame per byte) next to ordinary streams neither overflows the batch (which would fail every
    stream in it) nor keeps the others from making progress; a stream with more pending bytes than a batch holds is taken
    piecewise."""
    g = Gateway(device=0, max_conns=8, bytes_per_batch=1 << 16)
    try:
        evil = g.proxy_stream()
        good = [g.stream_chat_completions(R) for _ in range(3)]
        ok_body = (b'data: {"choices":[{"index":0,"delta":{"content":"hi"},"finish_reason":null}]}\n\n'
                   b'data: {"choices":[{"index":0,"delta":{},"finish_reason":"stop"}]}\n\n')
        flood = b"\n" * 100000                              # > bytes_per_batch: the pump takes what fits, round after round
        sent = 0
        got_evil = 0
        good_out = {s: [] for s in good}
        for s in good:
            assert g.upstream_write(s, ok_body) == len(ok_body)
        for _ in range(400):
            if sent < len(flood):
                sent += g.upstream_write(evil, flood[sent:sent + 30000])
            g.pump()
            while (e := g.proxy_step(evil)) is not None:
                assert e == b"\n"; got_evil += 1
            for s in good:
                try:
                    while (e := g.agent_recv(s)) is not None:
                        good_out[s].append(e)
                except EOFError:
                    pass
            if got_evil == len(flood):
                break
        assert got_evil == len(flood)
        for s in good:
            assert len(good_out[s]) == 3 and good_out[s][-1] == b"data: [DONE]\n\n"
    finally:
        g.close()
