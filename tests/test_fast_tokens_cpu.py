"""The decode kernel's whole-token shortcuts (csrc/sse_fast.h) against a table walk, on the CPU.

The shortcut phases are compiled from the same header by nvcc (into the decode kernel) and by g++ (into
tests/fast_tokens_check.cpp, a host model of one lane of the automaton). The check walks every payload of the
synthetic workloads, the reference fixtures and a few thousand mutations of them twice -- transitions only, and with the
shortcuts -- at all 16 window alignments and demands identical event logs."""
import json
import os
import re
import subprocess
import tempfile

import pytest

from inference_gateway_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "inference_gateway_b200", "csrc")


def _schema_names():
    src = open(os.path.join(CSRC, "sse_common.cuh")).read()
    names = re.findall(r'FD\("([a-z_]+)"', src)
    assert len(names) == 39
    return names


def _payloads():
    out = []
    for cfg in ("C2", "C3", "C4"):
        streams, _ = synth.make_config(cfg, n_streams=24)
        for body, _, _ in streams:
            for line in body.split(b"\n"):
                line = line.strip()
                if line.startswith(b"data: "):
                    out.append(line[6:])
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_fixtures.json")))

    def walk(o):
        if isinstance(o, str):
            for line in o.split("\n"):
                if line.startswith("data: "):
                    out.append(line[6:].encode())
        elif isinstance(o, dict):
            for v in o.values():
                walk(v)
        elif isinstance(o, list):
            for v in o:
                walk(v)
    walk(gold)
    # shapes the shortcuts must decline or bound correctly
    out += [
        b'{"id":"x","usage":{"prompt_tokens":12,"completion_tokens":0,"total_tokens":123456789012345678}}',
        b'{"usage":{"prompt_tokens":1234567890123456789,"completion_tokens":01,"total_tokens":-5}}',
        b'{"choices":[{"index":0,"delta":{"content":null,"tool_calls":[{"index":1,"id":"c","type":"function","function":{"name":"f","arguments":"{}"}}]},"logprobs":null,"finish_reason":"tool_calls"}]}',
        b'{"ID":"x","Choices":[],"\\u0069d":"y","id" :"z", "model": "m","created":1.5e3,"system_fingerprint":"fp","reasoning_content":"r"}',
        b'{"choices":[{"delta":{"content":"a"},"logprobs":{"content":[{"token":"t","logprob":-0.5,"bytes":[1,2,30],"top_logprobs":[]}]},"finish_reason":"stop"}],"x":true,"y":false,"z":nul}',
        b'{"a":tru', b'{"a":fals', b'{"a":12', b'{"created":1748534843', b'{"created":1748534843}', b'[1,22,333,0,00]', b'{"finish_reason":null}', b'null', b'0', b'17', b'true',
        b'{"completion_tokens_details":{"reasoning_tokens":0},"prompt_tokens_details":null}',
    ]
    return [p for p in out if p and b"\n" not in p]


@pytest.fixture(scope="module", params=[1, 2, 3])
def checker(request):
    tmp = tempfile.mkdtemp(prefix="fasttok_")
    exe = os.path.join(tmp, "fast_tokens_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-DFAST_REP=%d" % request.param, "-o", exe,
                           os.path.join(ROOT, "tests", "fast_tokens_check.cpp")])
    names = os.path.join(tmp, "names.txt")
    open(names, "w").write("\n".join(_schema_names()) + "\n")
    corpus = os.path.join(tmp, "corpus.txt")
    open(corpus, "wb").write(b"\n".join(_payloads()) + b"\n")
    return exe, names, corpus, request.param


@pytest.mark.parametrize("seed", [1, 2])
def test_shortcuts_equal_table_walk(checker, seed):
    exe, names, corpus, rep = checker
    r = subprocess.run([exe, names, corpus, "6000", str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-4000:]
    st = json.loads(r.stdout)
    # the shortcuts must actually fire on the clean workload: nearly every key, and the integers / literals
    assert st["clean_fast_keys"] > 20 * st["clean_table_key_ends"], st
    if rep > 1:       # (with one token per round the table walk has consumed a value's first byte before the next round looks at it)
        assert st["clean_fast_values"] > 0.2 * st["clean_fast_keys"], st
