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


def _per_stream(res):
    """Total frame count / record count / emitted bytes per segment, following the extra runs of long segments."""
    ff, fc = res.segs["frame_first"].astype(np.int64), res.segs["frame_count"].astype(np.int64)
    rc = res.segs["rec_count"].astype(np.int64)
    flen = res.frames["len"].astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(flen)])
    nbytes = csum[ff + fc] - csum[ff]
    fcount, rcount = fc.copy(), rc.copy()
    nxt = res.segs["next"]
    for i in np.nonzero(nxt != A.NONE)[0]:
        j = int(nxt[i])
        while j != A.NONE:
            r = res.runs[j]
            f0, fn = int(r["frame_first"]), int(r["frame_count"])
            fcount[i] += fn; rcount[i] += int(r["rec_count"]); nbytes[i] += csum[f0 + fn] - csum[f0]
            j = int(r["next"])
    return fcount, rcount, nbytes


def _joined(res, i):
    return b"".join(res.seg_frames(i))


def test_passthrough_is_the_identity_on_complete_lines(eng, workload):
    """Mode P at full size: per stream, the frames are exactly the '\\n'-terminated lines of the input, in order."""
    slot, res = _run(eng, workload, A.MODE_P)
    try:
        fcount, _, nbytes = _per_stream(res)
        n_lines = np.fromiter((b.count(b"\n") for b in workload), dtype=np.int64, count=N)
        assert np.array_equal(fcount, n_lines)
        expect = np.fromiter((b.rfind(b"\n") + 1 for b in workload), dtype=np.int64, count=N)
        assert np.array_equal(nbytes, expect)                    # unterminated tails are held back, nothing else
        total = np.fromiter((len(b) for b in workload), dtype=np.int64, count=N)
        assert np.array_equal(res.segs["carry_len"].astype(np.int64), total - expect)
        # byte identity on a stride of streams, and a checksum of checksums over them
        cin = cout = 0
        for i in range(0, N, 257):
            got = _joined(res, i)
            assert got == workload[i][:int(expect[i])]
            cin += sum(workload[i][:int(expect[i])]); cout += sum(got)
        assert cin == cout
    finally:
        eng.release(slot)


def test_reframe_accounting_and_idempotence(eng, workload):
    """Mode R at full size: every stream terminates at its finish chunk; frames == the stream's events up to it (canonical
    input: "data: X\\n\\n" re-frames to itself); one JSON_OK record per frame; re-framing the output is the identity."""
    slot, res = _run(eng, workload, R)
    try:
        assert np.all(res.segs["flags"] & A.SEG_TERMINATED)
        fcount, rcount, nbytes = _per_stream(res)
        assert np.array_equal(fcount, rcount)                      # no swallowed [DONE] before the finish chunk
        ends = np.empty(N, dtype=np.int64)                         # end of the terminating event in every input stream
        n_ev = np.empty(N, dtype=np.int64)
        for i, b in enumerate(workload):
            k = b.find(b'"finish_reason":"')
            e = b.index(b"\n\n", k) + 2
            ends[i] = e
            n_ev[i] = b.count(b"\n\n", 0, e)
        assert np.array_equal(fcount, n_ev)
        assert np.array_equal(nbytes, ends)
        sample = list(range(0, N, 129))
        outputs = []
        rflags = res.recs["flags"]
        for i in sample:
            got = _joined(res, i)
            assert got == workload[i][:int(ends[i])]               # canonical "data: X\n\n" re-frames to itself
            outputs.append(got)
            recs = res.seg_recs(i)
            assert all(rflags[r] & A.F_JSON_OK for r in recs)
            assert [bool(rflags[r] & A.F_TERMINATES) for r in recs] == [False] * (len(recs) - 1) + [True]
    finally:
        eng.release(slot)
    # idempotence: the emitted frames are canonical SSE, so feeding them back yields the same bytes
    slot, res2 = _run(eng, outputs, R)
    try:
        for j, body in enumerate(outputs):
            assert _joined(res2, j) == body
    finally:
        eng.release(slot)
