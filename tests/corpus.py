"""JSON documents that exercise every decode rule, and a byte-level mutator (shared by CPU and GPU tests)."""

TRICKY = [
    b'{}', b'null', b' null ', b'[]', b'"x"', b'5', b'true', b'', b' ', b'{', b'}', b'{"a"}', b'{"a":}', b'{"a":1,}',
    b'[1,]', b'{"a":1}}', b'{"a":1} x', b'{"a":01}', b'{"a":-}', b'{"a":1.}', b'{"a":.5}', b'{"a":1e}', b'{"a":1e+}',
    b'{"a":+1}', b'{"a":tru}', b'{"a":nul}', b'{"a":"\\x"}', b'{"a":"\\u12"}', b'{"a":"\\u12g4"}', b'{"a":"\x01"}',
    b'{"a":"\t"}', b'{"a":"\x7f"}', b'{"a":"\xff\xfe"}', b'{"choices":null}', b'{"choices":[]}', b'{"choices":{}}',
    b'{"choices":"x"}', b'{"choices":[null]}', b'{"choices":[1]}', b'{"choices":[{}]}', b'{"choices":[{"delta":null}]}',
    b'{"choices":[{"delta":[]}]}', b'{"choices":[{"delta":{"content":null}}]}', b'{"choices":[{"delta":{"content":5}}]}',
    b'{"choices":[{"delta":{"content":"a","content":"b"}}]}', b'{"choices":[{"delta":{"content":"a","content":null}}]}',
    b'{"Choices":[{"DELTA":{"CONTENT":"x"},"Finish_Reason":"stop"}]}',
    b'{"choice\xc5\xbf":[{"delta":{"content":"long s"}}]}', b'{"to\xe2\x84\xaaen":1,"choices":[{"delta":{"tool_calls":null}}]}',
    b'{"ch\\u006fices":[{"delta":{"c\\u006Fntent":"esc key"}}]}', b'{"choices":[{"delta":{"content":"\\ud83d\\ude00 \\ud83d x \\ude00 \\ud800\\u0041"}}]}',
    b'{"choices":[{"delta":{"content":"bad \xc0\xaf \xed\xa0\x80 \xf4\x90\x80\x80 \xe2\x82"}}]}',
    b'{"choices":[{"delta":{"content":"\\u0000\\b\\f\\n\\r\\t\\/\\\\\\""}}]}',
    b'{"choices":[{"finish_reason":"st\\u006fp"}]}', b'{"choices":[{"finish_reason":"tool_calls"}]}',
    b'{"choices":[{"finish_reason":"length"}]}', b'{"choices":[{"finish_reason":"weird_reason_that_is_long_and_unknown"}]}',
    b'{"choices":[{"finish_reason":""}]}', b'{"choices":[{"finish_reason":null}]}', b'{"choices":[{"finish_reason":7}]}',
    b'{"choices":[{"index":"0"}]}', b'{"choices":[{"index":1.0}]}', b'{"choices":[{"index":1e2}]}', b'{"choices":[{"index":-0}]}',
    b'{"choices":[{"index":9223372036854775807}]}', b'{"choices":[{"index":9223372036854775808}]}',
    b'{"choices":[{"index":-9223372036854775808}]}', b'{"choices":[{"index":-9223372036854775809}]}',
    b'{"created":12345678901234567890}', b'{"created":null}', b'{"created":true}', b'{"id":5}', b'{"id":null}', b'{"model":{}}',
    b'{"usage":null}', b'{"usage":{}}', b'{"usage":[]}', b'{"usage":{"prompt_tokens":1,"completion_tokens":2,"total_tokens":3}}',
    b'{"usage":{"prompt_tokens":"1"}}', b'{"usage":{"prompt_tokens":1},"usage":{"total_tokens":9}}',
    b'{"usage":{"prompt_tokens":1},"usage":null}', b'{"usage":{"PROMPT_TOKENS":7}}',
    b'{"choices":[{"delta":{"tool_calls":[]}}]}', b'{"choices":[{"delta":{"tool_calls":[null]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{}]}}]}', b'{"choices":[{"delta":{"tool_calls":[{"id":""}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"id":null,"function":null}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"function":{}}]}}]}', b'{"choices":[{"delta":{"tool_calls":[{"function":{"name":"n"}}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"function":{"arguments":"{\\"a\\":1}"},"index":3,"type":"function","id":"c"}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"index":0,"function":{"name":"a"}},{"index":0,"function":{"arguments":"x"}},{"index":-1}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"index":0}],"tool_calls":[{"index":1},{"index":2}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"index":0,"id":"a"}],"tool_calls":null}}]}',
    b'{"choices":[{"delta":{"tool_calls":{"index":0}}}]}', b'{"choices":[{"delta":{"tool_calls":[5]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"google":{"thought_signature":"s","other":[1,{"a":2}]}}}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"google":{"thought_signature":5}}}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"google":{"thought_signature":5,"thought_signature":"ok"}}}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"google":{"Thought_Signature":5}}}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"google":[1]}}]}}]}', b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"google":null}}]}}]}',
    b'{"choices":[{"delta":{"tool_calls":[{"extra_content":{"GOOGLE":"x"}}]}}]}',
    b'{"choices":[{"logprobs":{"content":[{"token":"a","logprob":-0.5,"bytes":[97],"top_logprobs":[{"token":"b","logprob":-9999.0,"bytes":null}]}],"refusal":null}}]}',
    b'{"choices":[{"logprobs":{"content":[{"logprob":3.4028235e38}]}}]}', b'{"choices":[{"logprobs":{"content":[{"logprob":3.4028236e38}]}}]}',
    b'{"choices":[{"logprobs":{"content":[{"logprob":340282356779733661637539395458142568447}]}}]}',
    b'{"choices":[{"logprobs":{"content":[{"logprob":340282356779733661637539395458142568448}]}}]}',
    b'{"choices":[{"logprobs":{"content":[{"logprob":-1e39}]}}]}', b'{"choices":[{"logprobs":{"content":[{"logprob":1e-400}]}}]}',
    b'{"choices":[{"logprobs":{"content":[{"logprob":0.00000e999}]}}]}', b'{"choices":[{"logprobs":{"content":[{"logprob":"x"}]}}]}',
    b'{"choices":[{"logprobs":{"content":[{"bytes":[1.5]}]}}]}', b'{"choices":[{"logprobs":{"content":{}}}]}',
    b'{"choices":[{"delta":{"content":"first"}},{"delta":{"content":"second","tool_calls":[{"id":"x"}]},"finish_reason":"stop"}]}',
    b'{"choices":[{"delta":{"content":"a"}}],"choices":[{"finish_reason":"stop"}]}',
    b'{"choices":[{"delta":{"content":"a"}}],"choices":null}', b'{"choices":[{"delta":{"content":"a"}}],"choices":[]}',
    b'{"unknown":{"deep":[1,2,{"x":[true,false,null,"s",1.5e-3]}]},"choices":[{"delta":{"content":"ok"}}]}',
    b'\t\r\n {"choices" : [ { "delta" : { "content" : "ws" } , "finish_reason" : null } ] } \r',
    b"[" * 128 + b"]" * 128, b"[" * 129 + b"]" * 129, b'{"a":' * 100 + b"1" + b"}" * 100,
    b'{"choices":[{"delta":{"role":"assistant","content":null,"reasoning":5}}]}', b'{"system_fingerprint":false}',
    b'{"reasoning_format":"raw","choices":[{"delta":{"reasoning_content":"r","refusal":null,"content":"c"}}]}',
]


def mutate(rng, doc: bytes) -> bytes:
    b = bytearray(doc)
    if not b:
        return bytes(b)
    for _ in range(int(rng.integers(1, 4))):
        op = int(rng.integers(0, 6))
        i = int(rng.integers(0, len(b)))
        if op == 0:
            b[i] = int(rng.integers(0x20, 0x7F))
        elif op == 1:
            del b[i]
        elif op == 2:
            alphabet = b'{}[]:,"\\0123456789.eE-+ tfn'
            b.insert(i, alphabet[int(rng.integers(0, len(alphabet)))])
        elif op == 3 and len(b) > 2:
            j = int(rng.integers(0, len(b)))
            b[i], b[j] = b[j], b[i]
        elif op == 4:
            b[i] = int(rng.integers(0x80, 0x100))
        else:
            b[i:i] = b'\\u00' + b'%02x' % int(rng.integers(0, 256))
        if not b:
            break
    return bytes(b).replace(b"\n", b" ")


