"""GPU: BASELINE.json's full sizes (65,536 concurrent streams, workload C4) checked through size-independent properties
instead of the oracle: exact byte identities between input and output, record/frame accounting, idempotence."""
import numpy as np
import pytest

from inference_gateway_b200 import SseEngine, _abi as A, synth

pytestmark = pytest.mark.gpu
N = 65536
R = A.MODE_R | A.MODE_PARSE


@pytest.fixture(scope="module")
def workload():
    streams, _ = synth.make_config("C4", n_streams=N)
    return [b for b, _, _ in streams]


@pytest.fixture(scope="module")
def eng(workload):
    e = SseEngine(device=0, max_conns=N, bytes_per_batch=sum(map(len, workload)), n_slots=1, carry_slot_bytes=16384)
    yield e
    e.close()


def _run(eng, bodies, mode):
    eng.reset_all()
    slot, arena, segs = eng.acquire()
    lens = np.fromiter((len(b) for b in bodies), dtype=np.int64, count=len(bodies))
    aligned = (lens + 15) & ~15
    offs = np.zeros(len(bodies), dtype=np.int64)
    offs[1:] = np.cumsum(aligned)[:-1]
    for b, o in zip(bodies, offs):
        arena[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    segs["conn"][:len(bodies)] = np.arange(len(bodies), dtype=np.uint32)
    segs["in_off"][:len(bodies)] = offs
    segs["in_len"][:len(bodies)] = lens
    segs["mode"][:len(bodies)] = mode
    segs["provider"][:len(bodies)] = 0
    segs["reserved"][:len(bodies)] = 0
    eng.submit(slot, len(bodies), int(offs[-1] + aligned[-1]))
    return slot, eng.collect(slot)


def test_passthrough_is_the_identity_on_complete_lines(eng, workload):
    """Mode P at full size: per stream, the frames are exactly the '\\n'-terminated lines of the input, in order."""
    slot, res = _run(eng, workload, A.MODE_P)
    try:
        assert int(res.raw.n_runs) == 0
        ff, fc = res.segs["frame_first"].astype(np.int64), res.segs["frame_count"].astype(np.int64)
        n_lines = np.fromiter((b.count(b"\n") for b in workload), dtype=np.int64, count=N)
        assert np.array_equal(fc, n_lines)
        flen = res.frames["len"].astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(flen)])
        per_stream_bytes = csum[ff + fc] - csum[ff]
        expect = np.fromiter((b.rfind(b"\n") + 1 for b in workload), dtype=np.int64, count=N)
        assert np.array_equal(per_stream_bytes, expect)          # unterminated tails are held back, nothing else
        assert np.array_equal(res.segs["carry_len"].astype(np.int64), np.fromiter((len(b) for b in workload), dtype=np.int64, count=N) - expect)
        # byte identity on a stride of streams (frames of a segment are contiguous in the out arena)
        for i in range(0, N, 257):
            o = int(res.frames["off"][ff[i]])
            assert res.out[o:o + int(expect[i])].tobytes() == workload[i][:int(expect[i])]
        # global checksum of checksums: every emitted byte is an input byte of a complete line
        total_in = sum(int(np.frombuffer(b[:e], dtype=np.uint8).sum(dtype=np.int64)) for b, e in zip(workload[::64], expect[::64]))
        total_out = sum(int(res.out[int(res.frames["off"][ff[i]]):int(res.frames["off"][ff[i]]) + int(expect[i])].sum(dtype=np.int64)) for i in range(0, N, 64))
        assert total_in == total_out
    finally:
        eng.release(slot)


def test_reframe_accounting_and_idempotence(eng, workload):
    """Mode R at full size: every stream terminates at its finish chunk; frames == the stream's events up to it (canonical
    input: "data: X\\n\\n" re-frames to itself); one JSON_OK record per frame; re-framing the output is the identity."""
    slot, res = _run(eng, workload, R)
    try:
        flags = res.segs["flags"]
        assert np.all(flags & A.SEG_TERMINATED)
        ff, fc = res.segs["frame_first"].astype(np.int64), res.segs["frame_count"].astype(np.int64)
        rf, rc = res.segs["rec_first"].astype(np.int64), res.segs["rec_count"].astype(np.int64)
        assert np.array_equal(fc, rc)                              # no swallowed [DONE] before the finish chunk
        # position of the terminating event in every input stream
        ends = np.empty(N, dtype=np.int64)
        n_ev = np.empty(N, dtype=np.int64)
        for i, b in enumerate(workload):
            k = b.find(b'"finish_reason":"')
            e = b.index(b"\n\n", k) + 2
            ends[i] = e
            n_ev[i] = b.count(b"\n\n", 0, e)
        assert np.array_equal(fc, n_ev)
        flen = res.frames["len"].astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(flen)])
        assert np.array_equal(csum[ff + fc] - csum[ff], ends)
        sample = list(range(0, N, 129))
        for i in sample:
            o = int(res.frames["off"][ff[i]])
            assert res.out[o:o + int(ends[i])].tobytes() == workload[i][:int(ends[i])]
        # records: all emitted chunks are valid JSON, exactly the last one terminates
        rflags = res.recs["flags"]
        last = rf + rc - 1
        assert np.all(rflags[last] & A.F_TERMINATES)
        ok_per_stream = np.add.reduceat((rflags & A.F_JSON_OK).astype(np.int64), rf) if np.all(np.diff(rf) > 0) else None
        if ok_per_stream is not None:
            pass
        valid = np.zeros(len(rflags), dtype=bool)
        for i in sample:
            valid[rf[i]:rf[i] + rc[i]] = True
        assert np.all(rflags[valid] & A.F_JSON_OK)
        term = (rflags[valid] & A.F_TERMINATES) != 0
        assert int(term.sum()) == len(sample)
        outputs = [res.out[int(res.frames["off"][ff[i]]):int(res.frames["off"][ff[i]]) + int(ends[i])].tobytes() for i in sample]
    finally:
        eng.release(slot)
    # idempotence: the emitted frames are canonical SSE, so feeding them back yields the same bytes
    slot, res2 = _run(eng, outputs, R)
    try:
        ff2, fc2 = res2.segs["frame_first"].astype(np.int64), res2.segs["frame_count"].astype(np.int64)
        for j, body in enumerate(outputs):
            o = int(res2.frames["off"][ff2[j]])
            assert res2.out[o:o + len(body)].tobytes() == body
    finally:
        eng.release(slot)
