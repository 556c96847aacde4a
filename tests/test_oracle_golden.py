"""CPU: pins the oracle against SURVEY.md Appendix B and the reference's own fixtures (semantic assertions the
reference tests make, transcribed in tests/golden/ref_fixtures.json by tools/make_golden.py)."""
import json
import os

import pytest

from oracle import orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.json")))["fixtures"]


def test_appendix_b_passthrough():
    assert orc.passthrough(b'data: {"a":1}\n\ndata: [DONE]\n\n').out == b'data: {"a":1}\n\ndata: [DONE]\n\n'     # P1
    v = orc.passthrough(b'data: {"a":1}\n\ndata: [DONE]\n\n')
    assert [l.out for l in v.lines] == [b'data: {"a":1}\n', b"\n", b"data: [DONE]\n", b"\n"]
    v = orc.passthrough(b"data: x\r\n\r\ndata: y")                                                                 # P2
    assert v.out == b"data: x\r\n\r\n" and v.tail_len == 7
    assert orc.passthrough(b"").out == b""                                                                          # P3
    assert orc.passthrough(b": ping\n\nevent: x\ndata: 1\n\n").out == b": ping\n\nevent: x\ndata: 1\n\n"          # P4
    big = b"x" * 10000 + b"\n"
    assert orc.passthrough(big).out == big                                                                          # P5


def test_appendix_b_reframe():
    r1 = b'data: {"choices":[{"index":0,"delta":{"content":"Hi"},"finish_reason":null}]}\n'
    assert orc.reframe(r1).out == r1 + b"\n"                                                                        # R1
    assert orc.reframe(b'\n: ping\nevent: m\ndata:{"x":1}\ndata: \n').out == b""                                    # R2
    assert orc.reframe(b'  data: {"x":1} \r\n').out == b'data: {"x":1}\n\n'                                         # R3
    assert orc.reframe(b'data:  {"x":1}\n').out == b'data:  {"x":1}\n\n'                                            # R4
    assert orc.reframe(b"data: [DONE]\n").out == b""                                                                # R5
    assert orc.reframe(b'data: {"choices":[{"index":0,"delta":{"content":"say [DONE]"},"finish_reason":null}]}\n').out == b""  # R6
    assert orc.reframe(b"data: {oops\n").out == b"data: {oops\n\n"                                                  # R7
    v = orc.reframe(b'data: {"choices":[{"index":0,"delta":{},"finish_reason":"stop"}]}\n\ndata: {"choices":[]}\n\ndata: [DONE]\n\n', True)
    assert v.out == b'data: {"choices":[{"index":0,"delta":{},"finish_reason":"stop"}]}\n\ndata: [DONE]\n\n'       # R8 + R9
    assert v.terminated and [l.kind for l in v.lines][1:] == [orc.L_UNREAD] * 5


def _bodies(fx):
    for it in fx["iterations"]:
        if fx["kind"] == "channel_elements":
            yield ("\n".join(it) + "\n").encode()
        else:
            yield (it[0] if it[0].endswith("\n") else it[0] + "\n").encode()


@pytest.mark.parametrize("fx", [f for f in GOLD if "content" in f["expect"]], ids=lambda f: f["name"])
def test_reference_fixture_semantics(fx):
    exp = fx["expect"]
    fin = {orc.FIN_STOP: "stop", orc.FIN_TOOL_CALLS: "tool_calls"}
    n_done = 0
    for i, body in enumerate(_bodies(fx)):
        last = i == len(fx["iterations"]) - 1
        v = orc.reframe(body, append_done=last)
        n_done += v.out.count(b"data: [DONE]\n\n")
        assert v.acc_content.decode() == exp["content"][i]                 # mcp_agent_test.go:569-597
        assert v.terminated and fin[v.term_finish] == exp["finish"][i]
        calls = orc.parse_tool_calls(v.builder) if v.has_tool_calls else []
        got = [dict(id=c["id"].decode(), name=c["name"].decode(), args=c["args"].decode()) for c in calls]
        assert got == exp["tool_calls"][i]                                  # mcp_agent_test.go:745-750
        assert v.has_tool_calls == bool(exp["tool_calls"][i])
        if exp["usage"][i] is not None:
            us = [l.chunk.usage for l in v.lines if l.chunk is not None and l.chunk.json_ok and l.chunk.usage]
            assert us and list(us[-1]) == exp["usage"][i]                   # usage rides on the finish chunk
    assert n_done == exp["done_frames"]                                     # middlewares/mcp_test.go:910


@pytest.mark.parametrize("fx", [f for f in GOLD if "parsed" in f["expect"]], ids=lambda f: f["name"])
def test_reference_fixture_builder(fx):
    calls = orc.parse_tool_calls(fx["iterations"][0][0].encode())
    got = [dict(id=c["id"].decode(), name=c["name"].decode(), args=c["args"].decode()) for c in calls]
    assert got == fx["expect"]["parsed"]


@pytest.mark.parametrize("fx", [f for f in GOLD if "frames" in f["expect"]], ids=lambda f: f["name"])
def test_reference_fixture_dropped_elements(fx):
    """Channel elements that are not "data: "-prefixed never reach the client (agent.go:186-188)."""
    for body in _bodies(fx):
        v = orc.reframe(body, append_done=True)
        assert [l.kind for l in v.lines] == [orc.L_DROPPED] * len(v.lines)
        assert v.out == b"data: [DONE]\n\n" * fx["expect"]["done_frames"]
        assert v.terminated == fx["expect"]["terminated"] and v.acc_content.decode() == fx["expect"]["acc_content"]


def test_every_sse_literal_of_the_reference_tests_is_a_fixture():
    """tools/make_golden.py walks every *_test.go of the reference for string literals that hold SSE bytes and records the ones
    no fixture contains: none."""
    audit = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.json")))["audit"]
    assert audit["n_literals"] >= 70 and audit["uncovered"] == []
    assert audit["files"] == ["tests/mcp_agent_test.go", "tests/middlewares/mcp_test.go"]


def test_telemetry_quirks():
    # only the last 4 "\n\n" pieces are searched for usage (telemetry.go:195-198)
    u = b'data: {"usage":{"prompt_tokens":1,"completion_tokens":2,"total_tokens":3}}\n\n'
    pad = b'data: {"choices":[]}\n\n'
    assert orc.telemetry(u + pad + b"data: [DONE]\n\n")[0] == (1, 2, 3)        # pieces: u, pad, DONE, ""
    assert orc.telemetry(u + pad * 2 + b"data: [DONE]\n\n")[0] == (0, 0, 0)    # 5 pieces: u is outside the last 4
    assert orc.telemetry(u.replace(b"\n\n", b"\r\n\r\n"))[0] == (1, 2, 3)       # single piece, trailing whitespace tolerated
    assert orc.telemetry((u + pad).replace(b"\n\n", b"\r\n\r\n"))[0] == (0, 0, 0)  # no "\n\n": one unparsable piece
    # tool calls without a name are dropped (telemetry.go:271)
    tc = b'data: {"choices":[{"delta":{"tool_calls":[{"index":0,"id":"a","function":{"arguments":"{}"}}]}}]}\n\n'
    assert orc.telemetry(tc)[1] == []
