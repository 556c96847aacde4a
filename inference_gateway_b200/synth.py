"""Deterministic synthetic SSE streams (SURVEY.md section 8d / BASELINE.md section 3).

Shapes are taken from the reference's own fixtures: full OpenAI envelope (tests/middlewares/mcp_test.go:485),
Groq role-first chunk with "content":null and tool_call with trailing "index" (:817, :822), envelope-less
chunks (tests/mcp_agent_test.go:529). anthropic / cohere / ollama are the same OpenAI-compatible shape with
provider-typical model strings, because the reference only ever talks to their compat endpoints
(providers/constants/constants.go:18-54). Seeded with 0xB200 + config index.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_WORDS = ("the of and to in is that for it as was with be by on not he this are or his from at which but have an "
          "had they you were their one all we can her has there been if more when will would who so no out up "
          "gateway stream token model provider latency batch kernel tensor memory warp shared async copy line "
          "chunk delta content tool call function argument result value number string object array").split()

MODELS = {
    "openai": "gpt-4o-mini-2024-07-18",
    "groq": "meta-llama/llama-4-scout-17b-instruct",
    "anthropic": "claude-sonnet-4-5-20250929",
    "cohere": "command-a-03-2025",
    "ollama": "llama3.2:3b",
    "bare": "",
}
ESCAPES = ['\\"', "\\\\", "\\n", "\\u00e9", "\\ud83d\\ude00", "\\t", "\\/"]
RAW_UTF8 = ["é", "ü", "漢", "😀"]


def _blob(rng: np.random.Generator, n: int = 1 << 18) -> str:
    idx = rng.integers(0, len(_WORDS), size=n // 4)
    return " ".join(_WORDS[i] for i in idx)


@dataclass
class StreamSpec:
    flavour: str = "openai"
    mean_bytes: int = 256
    n_content: int = 5
    tool_run: int = 0          # number of tool-call chunks (0: none)
    usage_chunk: bool = True   # trailing usage-only chunk (choices: [])
    sigma: float = 0.35
    crlf: bool = False
    with_done: bool = True


class Synth:
    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)
        self.blob = _blob(self.rng)
        self.n = 0

    def _text(self, n: int) -> str:
        n = max(1, n)
        o = int(self.rng.integers(0, len(self.blob) - n - 1))
        t = self.blob[o:o + n]
        r = self.rng.random()
        if r < 0.05:
            e = ESCAPES[int(self.rng.integers(0, len(ESCAPES)))]
            k = int(self.rng.integers(0, len(t)))
            t = t[:k] + e + t[k:]
        elif r < 0.06:
            e = RAW_UTF8[int(self.rng.integers(0, len(RAW_UTF8)))]
            k = int(self.rng.integers(0, len(t)))
            t = t[:k] + e + t[k:]
        return t

    def _env(self, spec: StreamSpec, sid: str, created: int, choice: str, tail: str = "") -> str:
        fl = spec.flavour
        if fl == "bare":
            return '{"id":"%s","choices":[%s]%s}' % (sid, choice, tail)
        head = '{"id":"%s","object":"chat.completion.chunk","created":%d,"model":"%s",' % (sid, created, MODELS[fl])
        if fl in ("openai", "ollama"):
            head += '"system_fingerprint":"fp_%s",' % ("ollama" if fl == "ollama" else sid[-10:])
        return head + '"choices":[%s]%s}' % (choice, tail)

    def stream(self, spec: StreamSpec) -> tuple[bytes, int]:
        """Returns (bytes of the upstream body, number of SSE data events)."""
        self.n += 1
        fl = spec.flavour
        sid = {"anthropic": "msg_01", "groq": "chatcmpl-", "cohere": "", "ollama": "chatcmpl-", "openai": "chatcmpl-",
               "bare": "test"}[fl] + "%012x" % int(self.rng.integers(0, 1 << 48))
        created = 1748534842 + self.n
        sep = "\r\n\r\n" if spec.crlf else "\n\n"
        ev = []
        null_content = ',"content":null' if fl == "groq" else ',"content":""'
        ev.append(self._env(spec, sid, created, '{"index":0,"delta":{"role":"assistant"%s},"finish_reason":null}' % null_content))
        extra = ',"logprobs":null' if fl == "openai" else ""
        utail = ',"usage":null' if fl == "anthropic" else ""
        content_slots = len(ev)
        ev.extend([None] * spec.n_content)
        fin = "stop"
        if spec.tool_run > 0:
            fin = "tool_calls"
            name = "mcp_" + _WORDS[int(self.rng.integers(0, len(_WORDS)))] + "_tool"
            cid = "call_%08x" % int(self.rng.integers(0, 1 << 32))
            args = '{"query":"%s","limit":%d,"tags":["a","b"]}' % (self._text(40 + 24 * spec.tool_run).replace("\\", ""), int(self.rng.integers(1, 100)))
            esc = args.replace("\\", "\\\\").replace('"', '\\"')
            # split the escaped argument text into tool_run-1 fragments at safe positions (never inside an escape)
            nfrag = max(1, spec.tool_run - 1)
            cuts = sorted(set(int(c) for c in self.rng.integers(1, max(2, len(esc) - 1), size=nfrag - 1)))
            frags, last = [], 0
            for c in cuts + [len(esc)]:
                while c < len(esc) and c > 0 and esc[c - 1] == "\\":
                    c += 1
                if c > last:
                    frags.append(esc[last:c]); last = c
            if fl == "groq":
                first = '{"id":"%s","type":"function","function":{"name":"%s","arguments":""},"index":0}' % (cid, name)
            else:
                first = '{"index":0,"id":"%s","type":"function","function":{"name":"%s","arguments":""}}' % (cid, name)
            ev.append(self._env(spec, sid, created, '{"index":0,"delta":{"tool_calls":[%s]},"finish_reason":null}' % first, utail))
            for fr in frags:
                tc = '{"index":0,"function":{"arguments":"%s"}}' % fr
                ev.append(self._env(spec, sid, created, '{"index":0,"delta":{"tool_calls":[%s]},"finish_reason":null}' % tc, utail))
        usage = '"usage":{"prompt_tokens":%d,"completion_tokens":%d,"total_tokens":%d}' % (
            int(self.rng.integers(5, 4000)), spec.n_content + spec.tool_run, 0)
        if fl == "groq":
            tail = ',"x_groq":{"id":"req_%s","usage":{"queue_time":0.0123,"prompt_tokens":12,"total_time":1.5e-1}}' % sid[-8:]
        elif spec.usage_chunk:
            tail = utail
        else:
            tail = "," + usage
        ev.append(self._env(spec, sid, created, '{"index":0,"delta":{},"finish_reason":"%s"}' % fin, tail))
        if spec.usage_chunk:
            ev.append(self._env(spec, sid, created, "", "," + usage).replace('"choices":[]', '"choices":[]'))
        # content deltas: sized so that the mean over the stream's data events is spec.mean_bytes
        fixed = sum(len(e) + 6 + len(sep) for e in ev if e is not None)
        k_ev = len(ev)
        base = len(self._env(spec, sid, created, '{"index":0,"delta":{"content":""}%s,"finish_reason":null}' % extra, utail)) + 6 + len(sep)
        if spec.n_content:
            m_c = max(96.0, (spec.mean_bytes * k_ev - fixed) / spec.n_content)
            mu = np.log(m_c) - spec.sigma ** 2 / 2
            for i in range(spec.n_content):
                tgt = int(min(4096, max(96, self.rng.lognormal(mu, spec.sigma))))
                txt = self._text(tgt - base)
                ev[content_slots + i] = self._env(
                    spec, sid, created, '{"index":0,"delta":{"content":"%s"}%s,"finish_reason":null}' % (txt, extra), utail)
        body = "".join("data: " + e + sep for e in ev)
        n_ev = len(ev)
        if spec.with_done:
            body += "data: [DONE]" + sep
            n_ev += 1
        return body.encode("utf-8"), n_ev


CONFIGS = {
    # name: (seed index, n_streams, mean bytes, flavours, mode bits, fraction of streams with a tool run)
    "C1": (0, 1, 256, ("ollama",), None, 0.0),
    "C2": (1, 4096, 256, ("openai",), 0, 0.0),
    "C3": (2, 16384, 512, ("anthropic",), 1 | 2, 0.0),
    "C4": (3, 65536, 512, ("cohere", "groq", "anthropic", "ollama"), 1 | 2, 0.10),
}


def make_config(name: str, n_streams: int | None = None, n_content: int | None = None, shard: int = 0):
    """Returns (list of (bytes, n_events, flavour), mode) for a BASELINE.json config.
    shard > 0 derives an independent stream population for another GPU (seed + 0x1000 * shard)."""
    idx, n, mean, flavours, mode, tool_frac = CONFIGS[name]
    if n_streams is not None:
        n = n_streams
    g = Synth(0xB200 + idx + 0x1000 * shard)
    out = []
    for i in range(n):
        fl = flavours[i % len(flavours)]
        tool = int(g.rng.integers(6, 13)) if g.rng.random() < tool_frac else 0
        nc = n_content if n_content is not None else (126 if name == "C1" else (7 if not tool else 2))
        spec = StreamSpec(flavour=fl, mean_bytes=mean, n_content=nc, tool_run=tool,
                          usage_chunk=fl not in ("cohere", "groq"))
        body, n_ev = g.stream(spec)
        out.append((body, n_ev, fl))
    return out, mode


def random_cuts(rng: np.random.Generator, data: bytes, n_parts: int) -> list[bytes]:
    """Simulated TCP segmentation: cut at seeded random byte positions."""
    if n_parts <= 1 or len(data) < 2:
        return [data]
    cuts = sorted(set(int(c) for c in rng.integers(1, len(data), size=n_parts - 1)))
    parts, last = [], 0
    for c in cuts + [len(data)]:
        parts.append(data[last:c]); last = c
    return parts
