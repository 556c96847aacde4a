"""CPU: ssegw_mcp_writer_step (host code of libssegpu.so) against an independent model of the writer loop of
handleMCPStreamingRequest (api/middlewares/mcp.go:253-299): terminal frame rule and the upstream-error sniff, whose
json.Unmarshal(line[6:], &struct{ Error string `json:"error"` }) is modelled with tests/go_model.py (CPython json for the
syntax, Go's key matching and type rules written from the encoding/json documentation)."""
import ctypes as C

import numpy as np

from inference_gateway_b200 import _abi as A
from tests import go_model as gm
from tests.corpus import TRICKY, mutate


def model(frame: bytes):
    if frame == b"data: [DONE]\n\n":
        return 1, 0
    if not (frame.startswith(b"data: {") and b'"error"' in frame):
        return 0, 0
    ok, v = gm.parse(frame[6:])
    if not ok or not isinstance(v, gm.Obj):
        return 0, 0
    for k, val in v.pairs:
        if k == "error" or gm.fold(k) == "ERROR":
            if not (isinstance(val, str) or val is None):
                return 0, 0          # UnmarshalTypeError: err != nil, no WriteHeader
    return 0, 1


def product(L, frame: bytes):
    flag = C.c_int(-1)
    stop = L.ssegw_mcp_writer_step(frame, len(frame), C.byref(flag))
    return stop, flag.value


HAND = [
    b"data: [DONE]\n\n", b"data: [DONE]\n", b"data: [DONE]\n\n\n", b" data: [DONE]\n\n", b"",
    b'data: {"error": "upstream said no"}\n\n', b'data: {"error":"x"}', b'data: {"error":null}\n\n',
    b'data: {"error":{"message":"m","type":"t"}}\n\n', b'data: {"error":5}\n\n', b'data: {"error":true}\n\n', b'data: {"error":[]}\n\n',
    b'data: {"Error":"x"}\n\n', b'data: {"ERROR":1,"x":"error"}\n\n', b'data: {"eRRoR":"fine","x":"error"}\n\n',
    b'data: {"a":{"error":5}}\n\n', b'data: {"a":[{"error":{}}],"b":"c"}\n\n',
    b'data: {"error":"a","error":2}\n\n', b'data: {"error":2,"error":"a"}\n\n',
    b'data: {"\\u0065rror":5,"x":"error"}\n\n', b'data: {"\\u0045RROR":"s","x":"error"}\n\n', b'data: {"err\\u006fr":[],"x":"error"}\n\n',
    b'data: {"error":"bad\n\n', b'data: {"error":"x"}} \n\n', b'data: {"error":"x"},\n\n', b'data: {"error":"x"}  \t\r\n',
    b'data: {"error":"\x01"}\n\n', b'data: {"error":"\\q"}\n\n', b'data: {"error":"\\ud800"}\n\n', b'data: {"error":"\xff\xfe"}\n\n',
    b'data: {"x":"say \\"error\\" twice"}\n\n', b'data: {"x":"\\"error\\""}\n\n', b'data:{"error":"x"}\n\n', b'data:  {"error":"x"}\n\n',
    b'data: {"choices":[{"delta":{"content":"an \\"error\\" occurred"}}],"error":"e"}\n\n',
    b'data: {"error" : "spaced" , "k" : [ 1 , 2.5e+3 , -0 , true , false , null , {} , [] ] }\n\n',
    b'data: {"error":01}\n\n', b'data: {"error":"x",}\n\n', b'data: {,"error":"x"}\n\n', b'data: {"error"}\n\n',
    b'data: {"errors":5,"x":"error"}\n\n', b'data: {"erro":5,"x":"error"}\n\n', b'data: {"error\xc5\xbf":5,"x":"error"}\n\n',
]


def test_mcp_writer_step_matches_model():
    L = A.load()
    frames = list(HAND)
    rng = np.random.default_rng(77)
    base = [d for d in TRICKY if d.startswith(b"{")]
    for i in range(4000):
        d = base[int(rng.integers(len(base)))]
        if i % 2:
            d = mutate(rng, d)
        k = int(rng.integers(6))
        inj = [b'"error":"m",', b'"error":{"code":1},', b'"Error":null,', b'"x":"error",', b'"error":7,', b""][k]
        if d.startswith(b"{") and inj:
            d = b"{" + inj + d[1:]
        frames.append(b"data: " + d + b"\n\n")
    n503 = 0
    for f in frames:
        exp, got = model(f), product(L, f)
        assert got == exp, f
        n503 += exp[1]
    assert n503 > 100       # the corpus exercises the positive branch too
