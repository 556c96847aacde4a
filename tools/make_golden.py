#!/usr/bin/env python3
"""Extracts the SSE fixtures the reference's own tests hold for the streaming path and writes them, with
the semantic expectations those tests assert, to tests/golden/ref_fixtures.json.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tools/make_golden.py
The reference tests assert semantic properties, not bytes; the expectations below are
transcribed from the cited assertions. The inputs are copied verbatim from the Go raw-string literals.
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_fixtures.json")


def raw_strings(path, lo, hi):
    """Go raw string literals (`...`) that start between lines lo..hi (1-based, inclusive)."""
    text = open(os.path.join(REF, path), encoding="utf-8").read()
    out = []
    for m in re.finditer(r"`([^`]*)`", text):
        line = text.count("\n", 0, m.start()) + 1
        if lo <= line <= hi:
            out.append(m.group(1))
    return out


def main():
    fx = []
    # tests/mcp_agent_test.go:529-537 "no tool calls streaming": channel elements fed to RunWithStream
    els = [s for s in raw_strings("tests/mcp_agent_test.go", 529, 537)]
    fx.append(dict(name="agent_no_tool_calls", source="tests/mcp_agent_test.go:529-537", kind="channel_elements",
                   iterations=[els],
                   expect=dict(content=["Hello there!"],           # :569-597 concatenated delta.content
                               finish=["stop"], usage=[[10, 3, 13]], tool_calls=[[]], done_frames=1)))
    # tests/mcp_agent_test.go:674-699 and :706-735 two-iteration tool-call stream
    it1 = raw_strings("tests/mcp_agent_test.go", 674, 699)
    it2 = raw_strings("tests/mcp_agent_test.go", 706, 735)
    fx.append(dict(name="agent_two_iterations_tool_calls", source="tests/mcp_agent_test.go:674-735", kind="channel_elements",
                   iterations=[it1, it2],
                   expect=dict(content=["I'll use both tools to help you.", "Based on the tool results, both tools executed successfully!"],
                               finish=["tool_calls", "stop"], usage=[[15, 8, 23], [25, 12, 37]],
                               # :745, :749-750 ExecuteTools receives these two calls with these argument maps
                               tool_calls=[[dict(id="call_123", name="mcp_test_tool", args='{"param":"value"}'),
                                            dict(id="call_456", name="mcp_other_tool", args='{"action":"execute"}')], []],
                               done_frames=1)))
    # tests/middlewares/mcp_test.go:485-501 whole-body streaming responses (already "\n\n"-separated)
    bodies = raw_strings("tests/middlewares/mcp_test.go", 484, 502)
    fx.append(dict(name="middleware_body_tool_calls", source="tests/middlewares/mcp_test.go:485-491", kind="body",
                   iterations=[[bodies[0]]],
                   expect=dict(content=[""], finish=["tool_calls"], usage=[None],
                               tool_calls=[[dict(id="call_123", name="test_function", args='{"param":"value"}')]], done_frames=1)))
    fx.append(dict(name="middleware_body_content", source="tests/middlewares/mcp_test.go:497-501", kind="body",
                   iterations=[[bodies[1]]],
                   expect=dict(content=["Hello there!"], finish=["stop"], usage=[None], tool_calls=[[]], done_frames=1)))
    # tests/middlewares/mcp_test.go:740-752 envelope-less tool-call fragments (input of the vacuous TestParseStreamingToolCalls)
    frag = raw_strings("tests/middlewares/mcp_test.go", 738, 753)
    fx.append(dict(name="tool_call_fragments_single", source="tests/middlewares/mcp_test.go:740-743", kind="builder",
                   iterations=[[frag[0]]],
                   expect=dict(parsed=[dict(id="call_123", name="mcp_test_tool", args='{"arg1":"value1","arg2":42}')])))
    fx.append(dict(name="tool_call_fragments_multi", source="tests/middlewares/mcp_test.go:750-752", kind="builder",
                   iterations=[[frag[2]]],
                   expect=dict(parsed=[dict(id="call_1", name="tool_one", args='{"x":1}'),
                                       dict(id="call_2", name="tool_two", args='{"y":2}')])))
    # tests/middlewares/mcp_test.go:816-869 Groq-style three iterations ("data: " + chunk per channel element)
    groq = raw_strings("tests/middlewares/mcp_test.go", 814, 870)
    groq = [g for g in groq if g.startswith("{")]
    its, cur = [], []
    for g in groq:
        cur.append("data: " + g)
        if '"finish_reason":"tool_calls"' in g or '"finish_reason":"stop"' in g:
            cur.append("data: [DONE]")
            its.append(cur); cur = []
    fx.append(dict(name="groq_three_iterations", source="tests/middlewares/mcp_test.go:816-869", kind="channel_elements",
                   iterations=its,
                   expect=dict(content=["I'll get pizza info", "Let me get more", "Based on pizza info, Margherita, Pepperoni Hawaiian"],
                               finish=["tool_calls", "tool_calls", "stop"], usage=[None, None, None],
                               tool_calls=[[dict(id="call_vxw1", name="get-pizza-info", args="{}")],
                                           [dict(id="call_vxw2", name="get-pizza-info", args="{}")], []],
                               done_frames=1)))      # :910 exactly one [DONE]
    # tests/middlewares/mcp_test.go:537 a channel element WITHOUT the "data: " prefix: the agent drops it (agent.go:186-188),
    # the channel closes, nothing is forwarded and only the final [DONE] is written (agent.go:140-143)
    bare = raw_strings("tests/middlewares/mcp_test.go", 537, 537)
    fx.append(dict(name="prefixless_channel_element", source="tests/middlewares/mcp_test.go:537", kind="channel_elements",
                   iterations=[bare], expect=dict(frames=0, terminated=False, acc_content="", done_frames=1)))
    audit = audit_coverage(fx)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(dict(reference="inference-gateway v0.24.0 @ ebf5d0e", generated_by="tools/make_golden.py", fixtures=fx, audit=audit),
              open(OUT, "w"), indent=1, ensure_ascii=False)
    print("wrote", OUT, len(fx), "fixtures", [len(i) for f in fx for i in f["iterations"]])
    print("audit:", audit["n_literals"], "SSE literals in", audit["files"], "- uncovered:", audit["uncovered"])


def audit_coverage(fx):
    """Every Go string literal under the reference's *_test.go files that holds SSE bytes (starts with "data: " or
    is a JSON chunk handed to a stream channel) must appear in some fixture above: the list of all of them, with
    file:line, and the ones no fixture contains (must stay empty)."""
    have = "\n".join(el for f in fx for it in f["iterations"] for el in it)
    lits, files = [], set()
    for root, _, names in os.walk(REF):
        for nm in names:
            if not nm.endswith("_test.go"):
                continue
            path = os.path.join(root, nm)
            text = open(path, encoding="utf-8").read()
            rel = os.path.relpath(path, REF)
            for m in re.finditer(r"`([^`]*)`", text):
                lit = m.group(1)
                if "data: " not in lit and not (lit.startswith("{") and '"choices"' in lit and '"delta"' in lit):
                    continue
                line = text.count("\n", 0, m.start()) + 1
                files.add(rel)
                for piece in [p for p in lit.split("\n") if p.strip()]:
                    lits.append(dict(at=f"{rel}:{line}", covered=piece.strip() in have or ("data: " + piece.strip()) in have,
                                     head=piece.strip()[:60]))
    return dict(n_literals=len(lits), files=sorted(files), uncovered=[l for l in lits if not l["covered"]],
                where=sorted(set(l["at"] for l in lits)))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("the reference tree is not present; golden fixtures are committed under tests/golden/")
    main()
