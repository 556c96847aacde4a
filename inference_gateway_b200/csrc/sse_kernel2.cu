// sse_kernel2.cu -- stages 2-4 of the default pipeline (work-item sort, decode, finalize), and the fused v2 kernel.
//
// The table-driven automaton (v2_round / v2_action / v2_finish_line): one lane per line steps a pushdown automaton over the
// payload it reads through a 16-byte register window. Syntax, type compatibility and field capture follow json.Unmarshal into
// CreateChatCompletionStreamResponse; keys are matched by a case-folding trie walked alongside; rare cases (escaped keys,
// int/float range checks) fall to the sequential helpers of sse_common.cuh; strings that need unquoting are queued and decoded
// by the whole warp.
//
//   sse_bucket_{hist,scan,scatter}_kernel  counting sort of the produce stage's work items by (length, shape class)
//   sse_decode_kernel                      persistent, one CTA per SM: warps pull 32 sorted items and run the automaton
//   sse_finalize_kernel                    early termination (agent.go:235-242): cut a segment's runs after its terminating chunk
//   sse_stream_kernel_v2 (SSE_FLAG_KERNEL_V2)  the earlier fused design: every warp an autonomous producer/consumer
//       (stage + split + classify + serialize, ring of work items, the same automaton as consumer); kept for the tests
#include <cuda_runtime.h>
#include <stdint.h>
#include "sse_common.cuh"
#include "sse_tables.h"

namespace {

using namespace ssetab;

constexpr int V2_WARPS = 16;
constexpr int V2_BUF = 6144;       // line window per warp
constexpr int RING = 128;          // work items per warp
constexpr int SEGSLOTS = 8;        // segments in flight per warp
#ifndef SSE_ROUNDS
#define SSE_ROUNDS 2
#endif
#ifndef SSE_KSTEPS
#define SSE_KSTEPS 2
#endif
constexpr int ROUNDS = SSE_ROUNDS;   // rounds between refills / busy checks
#ifndef SSE_SKIPW
#define SSE_SKIPW 4
#endif
constexpr int SKIPW = SSE_SKIPW;     // 16-byte windows a lane may cross per step while inside a long string value
constexpr int KSTEPS = SSE_KSTEPS;   // plain automaton steps per round before the pending actions run

struct SegSlot {
    unsigned long long term;       // min over terminating lines of (rec << 32 | frame); ~0ull: none
    uint32_t seg, conn;
    int32_t pending;               // lines not yet decoded (+1 while the producer still owns the segment)
    uint32_t used;
    uint32_t rmode;
    uint32_t pad;
};
struct LaneScratch {               // cold per-lane state (usage ints, the tool-call element being assembled)
    int64_t u_prompt, u_completion, u_total, tc_index;
    uint32_t tc_flags, tc_dec;
    uint32_t id_off, id_len, type_off, type_len, name_off, name_len, args_off, args_len;
};
struct WarpSmem2 {
    alignas(16) uint8_t buf[V2_BUF + 16];
    LineEnt lt[LT_MAX];
    uint16_t done_pos[DONE_MAX];
    uint32_t done_cnt;
    uint32_t ring_head, ring_tail;
    uint32_t pad0;
    uint4 ring[RING];              // src, len, rec, frame | slot << 29 (frame < 2^29)
    SegSlot slots[SEGSLOTS];
    LaneScratch ls[32];
};
struct CtaSmem2 {
    DfaTables T;
    WarpSmem2 w[V2_WARPS];
};
static_assert(sizeof(DfaTables) % 16 == 8 || sizeof(DfaTables) % 4 == 0, "tables are copied as 32-bit words");
static_assert(sizeof(CtaSmem2) <= 227 * 1024, "shared memory budget");

// per-string flags (cleared outside strings) and per-line flags
constexpr uint32_t SF_ESC = 1, SF_HI = 2, SF_UPPER = 4, SF_BAD = 8, SF_STRMASK = 15;
constexpr uint32_t SF_SYN = 0x100, SF_TYPE = 0x200, SF_DEPTH = 0x400, SF_GBAD = 0x800, SF_USAGE = 0x1000,
                   SF_TCNONNIL = 0x2000, SF_TCOPEN = 0x4000, SF_TCVALID = 0x8000, SF_CDEC = 0x10000, SF_RMODE = 0x20000, SF_CBAD = 0x40000, SF_CSET = 0x80000, SF_DONELINE = 0x100000;

struct Lane {
    uint32_t p, pe;                // out-arena offsets of the payload being decoded
    uint4 win;                     // the 16 bytes containing p
    uint32_t st, depth, skip, sd, cur, km, slen, sf, choices_count, n_choices, finish;
    unsigned long long ct, ct1, sstk;   // container-type bit stack (1 = array), 128 levels
    uint32_t content_off, content_len, tc_count, tc_first, tc_prev;
    uint32_t rec, frame, slot, plen;
    bool busy;
};

// RO = false: the bytes were written by another warp of the SAME kernel (fused v2): L2 only. RO = true: the payload arenas are
// read-only for the whole decode kernel, so the window may be cached in L1 (SSE_LDWIN 1).
#ifndef SSE_LDWIN
#define SSE_LDWIN 1
#endif
template <bool RO>
__device__ __forceinline__ uint4 ldwin16(const uint8_t *base, uint32_t off) {
    const uint4 *p = reinterpret_cast<const uint4 *>(base + (off & ~15u));
    return (RO && SSE_LDWIN) ? __ldg(p) : __ldcg(p);
}
__device__ __forceinline__ bool lane_live(const Lane &L) {
    return L.sd >= 3 && ((L.sstk >> 10) & 31ull) == N_CHOICE && L.choices_count == 1;
}
__device__ __forceinline__ uint32_t lane_top(const Lane &L) { return (uint32_t)((L.sstk >> (5 * (L.sd - 1))) & 31ull); }
__device__ __forceinline__ void value_done(Lane &L) {
    if (L.depth == 0) { L.st = S_END; return; }
    const uint32_t d = L.depth - 1;
    const unsigned long long bits = d < 64 ? L.ct : L.ct1;
    L.st = ((bits >> (d & 63u)) & 1ull) ? (uint32_t)S_AFTA : (uint32_t)S_AFTO;
}

// Span of a captured string: without escapes it is the payload's own bytes (no call, nothing through local memory); the
// rare decoded case goes through capture() (text-arena allocation + queued warp-cooperative unquote).
__device__ __forceinline__ Span capture_v2(const KParams &P, LaneJobs *J, uint32_t s, uint32_t e, int dec, uint32_t *patch) {
    if (!dec) { Span r; r.off = s; r.len = e - s; r.text = false; return r; }
    ParseCtx cx; cx.jobs = J; cx.sm = P.out; cx.P = &P; cx.S = nullptr; cx.emitted = true; cx.out_delta = 0;
    return capture(cx, (int)s, (int)e, dec, patch);
}

__device__ void v2_flush_tc(const KParams &P, Lane &L, LaneScratch &S, LaneJobs *J) {
    if ((S.tc_flags & SSE_TC_HAS_ID) || ((S.tc_flags & SSE_TC_HAS_FUNC) && (S.name_len || S.args_len))) L.sf |= SF_TCVALID;
    L.sf &= ~SF_TCOPEN;
    uint32_t idx = atomicAdd(&P.ctr->n_tcs, 1u);
    if (idx >= P.cap_tcs) { sse_overflow(P.ctr, SSE_OVF_TCS); return; }
    sse_tc *rec = &P.tcs[idx];
    Span id = capture_v2(P, J, S.id_off, S.id_off + S.id_len, S.tc_dec & 3, &rec->id_len);
    Span ty = capture_v2(P, J, S.type_off, S.type_off + S.type_len, (S.tc_dec >> 2) & 3, &rec->type_len);
    Span nm = capture_v2(P, J, S.name_off, S.name_off + S.name_len, (S.tc_dec >> 4) & 3, &rec->name_len);
    Span ar = capture_v2(P, J, S.args_off, S.args_off + S.args_len, (S.tc_dec >> 6) & 3, &rec->args_len);
    sse_tc o;
    o.index = S.tc_index;
    o.flags = S.tc_flags | (id.text ? SSE_TC_ID_TEXT : 0) | (ty.text ? SSE_TC_TYPE_TEXT : 0) |
              (nm.text ? SSE_TC_NAME_TEXT : 0) | (ar.text ? SSE_TC_ARGS_TEXT : 0);
    o.next = SSE_NONE;
    o.id_off = id.off; o.id_len = id.len; o.type_off = ty.off; o.type_len = ty.len;
    o.name_off = nm.off; o.name_len = nm.len; o.args_off = ar.off; o.args_len = ar.len;
    *rec = o;            // queued unquote jobs overwrite the *_len fields when the warp drains them
    if (L.tc_first == SSE_NONE) L.tc_first = idx; else P.tcs[L.tc_prev].next = idx;
    L.tc_prev = idx;
}

__device__ void v2_elem_begin(const KParams &P, Lane &L, LaneScratch &S, LaneJobs *J) {
    if (L.skip > 0) { L.cur = TY_SKIP; return; }
    uint32_t nd = lane_top(L);
    if (nd == A_CHOICES) { L.choices_count++; L.cur = TY_STRUCT | (N_CHOICE << 4); }
    else if (nd == A_TOOLCALLS) {
        L.cur = TY_STRUCT | (N_TC << 4);
        if (lane_live(L)) {
            if (L.sf & SF_TCOPEN) v2_flush_tc(P, L, S, J);
            L.sf |= SF_TCOPEN; L.tc_count++;
            S.tc_index = 0; S.tc_flags = 0; S.tc_dec = 0;
            S.id_off = S.id_len = S.type_off = S.type_len = S.name_off = S.name_len = S.args_off = S.args_len = 0;
        }
    }
    else if (nd == A_TOKLP) L.cur = TY_STRUCT | (N_TOKLP << 4);
    else if (nd == A_TOPLP) L.cur = TY_STRUCT | (N_TOPLP << 4);
    else L.cur = TY_INT;
}

__device__ void v2_null(Lane &L, LaneScratch &S) {
    uint32_t ty = L.cur & 15u, tgt = (L.cur >> 9) & 15u;
    if (ty == TY_TS) { L.sf &= ~SF_GBAD; return; }
    switch (tgt) {
    case TG_CHOICES:
        L.n_choices = 0; L.choices_count = 0; L.finish = SSE_FIN_NONE; L.content_off = L.content_len = 0;
        L.sf &= ~(SF_CDEC | SF_CBAD | SF_CSET | SF_TCNONNIL | SF_TCOPEN | SF_TCVALID);
        L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE;
        break;
    case TG_USAGE: L.sf &= ~SF_USAGE; S.u_prompt = S.u_completion = S.u_total = 0; break;
    case TG_TOOLCALLS:
        if (lane_live(L)) { L.sf &= ~(SF_TCNONNIL | SF_TCOPEN | SF_TCVALID); L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE; }
        break;
    case TG_TC_ID: if ((L.sf & SF_TCOPEN) && lane_live(L)) { S.tc_flags &= ~SSE_TC_HAS_ID; S.id_off = S.id_len = 0; S.tc_dec &= ~3u; } break;
    case TG_TC_TYPE: if ((L.sf & SF_TCOPEN) && lane_live(L)) { S.tc_flags &= ~SSE_TC_HAS_TYPE; S.type_off = S.type_len = 0; S.tc_dec &= ~12u; } break;
    case TG_TC_FUNCTION:
        if ((L.sf & SF_TCOPEN) && lane_live(L)) { S.tc_flags &= ~SSE_TC_HAS_FUNC; S.name_off = S.name_len = S.args_off = S.args_len = 0; S.tc_dec &= ~0xF0u; }
        break;
    default: break;
    }
}

__device__ void v2_number_end(const KParams &P, Lane &L, LaneScratch &S, uint32_t end) {
    const uint32_t ty = L.cur & 15u, tgt = (L.cur >> 9) & 15u;
    const bool is_int = L.st == S_NZERO || L.st == S_NINT;
    const uint32_t start = end - L.slen - 1;
    if (ty == TY_INT) {
        if (!is_int) L.sf |= SF_TYPE;
        else if (end - start > 18 || tgt != TG_NONE) {
            int64_t v;
            if (!parse_i64(P.out, (int)start, (int)end, v)) L.sf |= SF_TYPE;
            else if (tgt == TG_PROMPT) S.u_prompt = v;
            else if (tgt == TG_COMPLETION) S.u_completion = v;
            else if (tgt == TG_TOTAL) S.u_total = v;
            else if (tgt == TG_TC_INDEX) { if ((L.sf & SF_TCOPEN) && lane_live(L)) S.tc_index = v; }
        }
    } else if (ty == TY_F32) { if (f32_overflows(P.out, (int)start, (int)end)) L.sf |= SF_TYPE; }
    else if (ty == TY_TS) L.sf |= SF_GBAD;
    else if (ty != TY_SKIP) L.sf |= SF_TYPE;
}

// returns true when the current byte has to be looked up again in the new state
__device__ bool v2_action(const KParams &P, const DfaTables &T, Lane &L, LaneScratch &S, LaneJobs *J, uint32_t t) {
    switch (t) {
    case A_OPEN_OBJ: case A_OPEN_ARR: {
        const bool arr = t == A_OPEN_ARR;
        if (L.depth >= 128) { L.sf |= SF_DEPTH | SF_SYN; L.p = L.pe - 1; L.st = S_END; return false; }
        if (L.depth < 64) L.ct = (L.ct & ~(1ull << L.depth)) | ((unsigned long long)arr << L.depth);
        else L.ct1 = (L.ct1 & ~(1ull << (L.depth - 64))) | ((unsigned long long)arr << (L.depth - 64));
        L.depth++;
        const uint32_t ty = L.cur & 15u;
        if (L.skip > 0 || ty == TY_SKIP) L.skip++;
        else {
            const uint32_t okmask = arr ? ((1u << TY_SLICE) | (1u << TY_PSLICE))
                                        : ((1u << TY_STRUCT) | (1u << TY_PSTRUCT) | (1u << TY_ROOT) | (1u << TY_GOOGLE));
            if (!((okmask >> ty) & 1u)) { L.sf |= (ty == TY_TS) ? SF_GBAD : SF_TYPE; L.skip++; }
            else {
                const uint32_t sub = (L.cur >> 4) & 31u, tgt = (L.cur >> 9) & 15u;
                const bool live = lane_live(L);
                L.sstk = (L.sstk & ~(31ull << (5 * L.sd))) | ((unsigned long long)sub << (5 * L.sd));
                L.sd++;
                if (tgt != TG_NONE) {
                    if (tgt == TG_USAGE) L.sf |= SF_USAGE;
                    else if (tgt == TG_TC_FUNCTION) { if (live && (L.sf & SF_TCOPEN)) S.tc_flags |= SSE_TC_HAS_FUNC; }
                    else if (tgt == TG_CHOICES) L.choices_count = 0;
                    else if (tgt == TG_TOOLCALLS) {
                        if (live) { L.sf = (L.sf | SF_TCNONNIL) & ~(SF_TCOPEN | SF_TCVALID); L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE; }
                    }
                }
                if (sub == N_GOOGLE) L.sf &= ~SF_GBAD;
            }
        }
        L.st = arr ? S_ARR0 : S_OBJ0;
        return false;
    }
    case A_CLOSE_OBJ: case A_CLOSE_ARR: {
        L.depth--;
        if (L.skip > 0) L.skip--;
        else {
            const uint32_t node = lane_top(L);
            L.sd--;
            if (node == A_CHOICES) L.n_choices = L.choices_count;
            else if (node == A_TOOLCALLS) { if (lane_live(L) && (L.sf & SF_TCOPEN)) v2_flush_tc(P, L, S, J); }
            else if (node == N_GOOGLE) { if (L.sf & SF_GBAD) L.sf |= SF_TYPE; }
        }
        value_done(L);
        return false;
    }
    case A_KEY_END: {
        uint32_t cur = TY_SKIP;
        if (L.skip == 0) {
            const uint32_t node = lane_top(L);
            if (L.sf & (SF_ESC | SF_HI)) {   // escaped / non-ASCII key: unquote and fold like encoding/json does
                uint8_t tmp[72];
                uint32_t n = json_unquote(P.out, (int)(L.p - L.slen), (int)L.p, tmp, 64);
                int f = (n <= 64) ? match_field(c_schema, (int)node, tmp, (int)n) : -1;
                if (f >= 0) cur = c_schema.f[f].ty | ((uint32_t)c_schema.f[f].sub << 4) | ((uint32_t)c_schema.f[f].tgt << 9);
            } else {
                const uint32_t name = T.accept[L.km];
                if (name != 0xFFu) {
                    uint32_t f = T.field[node * NNAMES + name];
                    if ((f & FIELD_VALID) && !(node == N_GOOGLE && (L.sf & SF_UPPER))) cur = f & 0x1FFFu;
                }
            }
        }
        L.cur = cur;
        L.st = S_COLON;
        return false;
    }
    case A_VSTR_END: {
        const uint32_t ty = L.cur & 15u, tgt = (L.cur >> 9) & 15u;
        if (ty == TY_STR || ty == TY_PSTR) {
            if (tgt != TG_NONE && lane_live(L)) {
                const uint32_t start = L.p - L.slen, len = L.slen;
                const uint32_t d2 = ((L.sf & SF_ESC) ? 1u : 0u) | ((L.sf & SF_BAD) ? 2u : 0u);
                switch (tgt) {
                case TG_CONTENT:
                    L.content_off = start; L.content_len = len;
                    L.sf = (L.sf & ~(SF_CDEC | SF_CBAD)) | SF_CSET | ((d2 & 1u) ? SF_CDEC : 0u) | ((d2 & 2u) ? SF_CBAD : 0u);
                    break;
                case TG_FINISH:
                    if (len == 0) L.finish = SSE_FIN_NONE;
                    else if (!d2) { uint32_t nm = T.accept[L.km]; uint32_t fv = nm != 0xFFu ? T.finmap[nm] : 0xFFu; L.finish = fv != 0xFFu ? fv : (uint32_t)SSE_FIN_OTHER; }
                    else {
                        uint8_t tmp[40];
                        uint32_t n = json_unquote(P.out, (int)start, (int)L.p, tmp, 32);
                        L.finish = (n <= 32) ? classify_finish(tmp, (int)n) : (uint32_t)SSE_FIN_OTHER;
                    }
                    break;
                case TG_TC_ID: if (L.sf & SF_TCOPEN) { S.tc_flags |= SSE_TC_HAS_ID; S.id_off = start; S.id_len = len; S.tc_dec = (S.tc_dec & ~3u) | d2; } break;
                case TG_TC_TYPE: if (L.sf & SF_TCOPEN) { S.tc_flags |= SSE_TC_HAS_TYPE; S.type_off = start; S.type_len = len; S.tc_dec = (S.tc_dec & ~12u) | (d2 << 2); } break;
                case TG_NAME: if (L.sf & SF_TCOPEN) { S.name_off = start; S.name_len = len; S.tc_dec = (S.tc_dec & ~0x30u) | (d2 << 4); } break;
                case TG_ARGS: if (L.sf & SF_TCOPEN) { S.args_off = start; S.args_len = len; S.tc_dec = (S.tc_dec & ~0xC0u) | (d2 << 6); } break;
                default: break;
                }
            }
        } else if (ty == TY_TS) L.sf &= ~SF_GBAD;
        else if (ty != TY_SKIP) L.sf |= SF_TYPE;
        value_done(L);
        return false;
    }
    case A_BAD_STAY: L.sf |= SF_BAD; L.st = S_VSTR; return false;
    case A_BAD_REDO: L.sf |= SF_BAD; L.st = S_VSTR; return true;
    case A_NUM_END: v2_number_end(P, L, S, L.p); value_done(L); return true;
    case A_LIT_TRUE: case A_LIT_FALSE: {
        const uint32_t ty = L.cur & 15u;
        if (ty == TY_TS) L.sf |= SF_GBAD; else if (ty != TY_SKIP) L.sf |= SF_TYPE;
        value_done(L);
        return false;
    }
    case A_LIT_NULL: v2_null(L, S); value_done(L); return false;
    case A_ELEM_REDO: v2_elem_begin(P, L, S, J); L.st = S_VAL; return true;
    case A_COMMA_ARR: v2_elem_begin(P, L, S, J); L.st = S_VAL; return false;
    default:   // A_ERR
        L.sf |= SF_SYN; L.p = L.pe - 1; L.st = S_END;
        return false;
    }
}

// A line retires: final syntax check, record, termination bookkeeping (agent.go:205-242).
__device__ bool v2_finish_line(const KParams &P, Lane &L, LaneScratch &S, LaneJobs *J) {
    bool terminates = false;
    if (!(L.sf & SF_SYN)) {
        if (L.depth == 0 && (L.st == S_NZERO || L.st == S_NINT || L.st == S_NFRAC || L.st == S_NEXP)) {
            v2_number_end(P, L, S, L.pe);
            L.st = S_END;
        }
        if (L.st != S_END) L.sf |= SF_SYN;
    }
    sse_rec r;
    r.frame = L.frame; r.flags = 0; r.content_off = r.content_len = 0; r.tc_first = SSE_NONE; r.tc_count = 0; r.n_choices = 0;
    r.usage = SSE_NONE;
    if (L.sf & SF_DEPTH) r.flags |= SSE_F_DEPTH_LIMIT;
    if (L.sf & SF_DONELINE) r.flags |= SSE_F_DONE_LINE;      // swallowed by the reframe, parsed for agent.go:377-402
    if (!(L.sf & (SF_SYN | SF_TYPE))) {
        r.flags |= SSE_F_JSON_OK;
        r.n_choices = (uint16_t)min(L.n_choices, 0xFFFFu);
        if (L.sf & SF_USAGE) {
            uint32_t idx = atomicAdd(&P.ctr->n_usages, 1u);
            if (idx < P.cap_usages) {
                sse_usage u; u.prompt_tokens = S.u_prompt; u.completion_tokens = S.u_completion; u.total_tokens = S.u_total;
                P.usages[idx] = u; r.usage = idx; r.flags |= SSE_F_HAS_USAGE;
            } else sse_overflow(P.ctr, SSE_OVF_USAGES);
        }
        if (L.n_choices > 0) {
            Span ct = capture_v2(P, J, L.content_off, L.content_off + L.content_len, ((L.sf & SF_CDEC) ? 1 : 0) | ((L.sf & SF_CBAD) ? 2 : 0),
                                 &P.recs[L.rec].content_len);
            r.content_off = ct.len ? ct.off : 0; r.content_len = ct.len;
            if (ct.text && ct.len) r.flags |= SSE_F_CONTENT_TEXT;
            r.flags |= L.finish << SSE_F_FINISH_SHIFT;
            if (L.sf & SF_TCNONNIL) r.flags |= SSE_F_TC_NONNIL;
            if (L.sf & SF_TCVALID) r.flags |= SSE_F_TC_VALID;
            r.tc_first = L.tc_count ? L.tc_first : SSE_NONE;
            r.tc_count = (uint16_t)min(L.tc_count, 0xFFFFu);
            if ((L.sf & SF_RMODE) && (L.finish == SSE_FIN_STOP || L.finish == SSE_FIN_TOOL_CALLS)) {
                r.flags |= SSE_F_TERMINATES;
                terminates = true;
            }
        }
    }
    r.payload_len = L.plen;
    P.recs[L.rec] = r;
    L.busy = false;
    return terminates;
}

// Called by whichever lane (or the producer) drops a segment's pending count to zero.
__device__ void v2_finalize_segment(const KParams &P, SegSlot &sl) {
    if (sl.term != ~0ull) {
        const uint32_t trec = (uint32_t)(sl.term >> 32), tframe = (uint32_t)sl.term;
        sse_seg_result r = P.seg_results[sl.seg];
        sse_run *run = &r.run;
        for (;;) {   // find the run that holds the terminating record; everything after it was never read by the reference
            if (trec >= run->rec_first && trec < run->rec_first + run->rec_count) {
                run->rec_count = trec - run->rec_first + 1;
                run->frame_count = tframe - run->frame_first + 1;
                run->next = SSE_NONE;
                break;
            }
            if (run->next == SSE_NONE) break;
            run = &P.runs[run->next];
        }
        r.flags |= SSE_SEG_TERMINATED;
        r.carry_len = 0;
        P.seg_results[sl.seg] = r;
        ConnState ns; ns.carry_len = 0; ns.flags = CONN_FINISHED;
        P.conns[sl.conn] = ns;
    }
    __threadfence_block();
    sl.used = 0;
}

// One round of the per-lane automaton: KSTEPS plain steps, then the pending action (if any) of every lane.
template <bool RO>
__device__ __forceinline__ void v2_round(const KParams &P, const DfaTables &T, Lane &L, LaneScratch &S, LaneJobs *J) {
    uint32_t pend = 0;                       // action | cls << 8 | in_str << 16 | in_tok << 17
    // phase A (once per round, only the lanes inside a long string value): jump to the next '"', '\\', control or non-ASCII
    // byte, up to SKIPW windows. Keeping it out of the step loop means a warp whose lanes are not all in the same phase
    // executes this path once per round, not once per step.
    if (L.p < L.pe && L.st == S_VSTR && L.km == TRIE_DEAD) {
        #pragma unroll 1
        for (int w = 0; w < SKIPW; w++) {
            const uint32_t i = L.p & 15u;
            const uint32_t s0 = special_mask4(L.win.x), s1 = special_mask4(L.win.y), s2 = special_mask4(L.win.z), s3 = special_mask4(L.win.w);
            if (i == 0 && (s0 | s1 | s2 | s3) == 0 && L.p + 16u <= L.pe) {      // a whole window of plain string bytes
                L.p += 16u; L.slen += 16u;
                if (L.p >= L.pe) break;
                L.win = ldwin16<RO>(P.out, L.p);
                continue;
            }
            unsigned long long lo = ((unsigned long long)s1 << 32) | s0;
            unsigned long long hi = ((unsigned long long)s3 << 32) | s2;
            if (i < 8) lo &= ~0ull << (i * 8); else { lo = 0; hi &= ~0ull << ((i - 8) * 8); }
            const uint32_t j = lo ? (uint32_t)(__ffsll((long long)lo) - 1) >> 3 : (hi ? 8u + ((uint32_t)(__ffsll((long long)hi) - 1) >> 3) : 16u);
            const uint32_t n = min(j - i, L.pe - L.p);
            if (n == 0) break;
            L.p += n; L.slen += n;
            if ((L.p & 15u) != 0 || L.p >= L.pe) break;     // stopped at a special byte or at the end of the payload
            L.win = ldwin16<RO>(P.out, L.p);
        }
    }
    // phase B: plain automaton steps
    #pragma unroll
    for (int k = 0; k < KSTEPS; k++) {
        if (L.p < L.pe && pend == 0) {
            const uint32_t wsel = (L.p >> 2) & 3u;
            const uint32_t w01 = (wsel & 1u) ? L.win.y : L.win.x, w23 = (wsel & 1u) ? L.win.w : L.win.z;
            const uint32_t w = (wsel & 2u) ? w23 : w01;
            const uint32_t c = (w >> ((L.p & 3u) * 8u)) & 0xFFu;
            const uint32_t e = T.clssym[c];
            const uint32_t cls = e & 63u;
            const bool in_str = L.st >= S_KSTR, in_tok = L.st >= S_NMINUS;
            const uint32_t t = T.tr[L.st * NCLS + cls];
            if (t < A_FIRST) {
                L.km = in_str ? (uint32_t)T.kt[L.km * NSYM + ((e >> 8) & 31u)] : (uint32_t)TRIE_ROOT;
                L.sf = in_str ? (L.sf | (e >> 13)) : (L.sf & ~SF_STRMASK);
                L.slen = in_tok ? L.slen + 1 : 0;
                L.st = t;
                L.p++;
                if ((L.p & 15u) == 0 && L.p < L.pe) L.win = ldwin16<RO>(P.out, L.p);
            } else pend = t | (cls << 8) | (in_str ? 0x10000u : 0u) | (in_tok ? 0x20000u : 0u);
        }
    }
    if (pend) {
        uint32_t t = pend & 0xFFu;
        const uint32_t cls = (pend >> 8) & 0xFFu;
        for (;;) {
            if (!v2_action(P, T, L, S, J, t)) break;         // the action chose the next state
            t = T.tr[L.st * NCLS + cls];                     // redo: same byte, new state
            if (t < A_FIRST) { L.st = t; break; }
        }
        const bool in_str = (pend & 0x10000u) != 0;
        L.km = in_str ? (uint32_t)TRIE_DEAD : (uint32_t)TRIE_ROOT;
        const uint32_t nf = (cls == C_BSLASH ? SF_ESC : 0u) | (cls >= C_H80 ? SF_HI : 0u);
        L.sf = in_str ? (L.sf | nf) : (L.sf & ~SF_STRMASK);
        L.slen = (pend & 0x20000u) ? L.slen + 1 : 0;
        L.p++;
        if ((L.p & 15u) == 0 && L.p < L.pe) L.win = ldwin16<RO>(P.out, L.p);
    }
}

struct Producer {                  // warp-uniform coroutine state of the segment being produced
    int phase;                     // 0 idle, 1 load window, 2 rounds, 3 done
    uint32_t s, conn, mode, slot;
    sse_seg seg;
    int consumed, pend, base, fill, pos, clen;
    bool in_long, dead, overflow, more;
    RunChain rc;
};

__global__ void __launch_bounds__(V2_WARPS * 32, 1)
sse_stream_kernel_v2(const __grid_constant__ KParams P, const DfaTables *__restrict__ gT) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaSmem2 &cs = *reinterpret_cast<CtaSmem2 *>(smem_raw);
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(gT);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&cs.T);
        for (int i = threadIdx.x; i < (int)(sizeof(DfaTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    WarpSmem2 &W = cs.w[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    if (lane == 0) { W.ring_head = W.ring_tail = 0; W.done_cnt = 0; }
    if (lane < SEGSLOTS) { W.slots[lane].used = 0; W.slots[lane].pending = 0; W.slots[lane].term = ~0ull; }
    __syncthreads();
    const DfaTables &T = cs.T;
    uint8_t *buf = W.buf;
    LaneScratch &S = W.ls[lane];

    Lane L; L.busy = false; L.p = L.pe = 0; L.win = make_uint4(0, 0, 0, 0);
    L.st = S_END; L.depth = L.skip = L.sd = 0; L.cur = 0; L.km = TRIE_ROOT; L.slen = 0; L.sf = 0; L.choices_count = L.n_choices = 0;
    L.finish = 0; L.ct = L.ct1 = L.sstk = 0; L.content_off = L.content_len = 0; L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE;
    L.rec = L.frame = L.slot = L.plen = 0;

    Producer pr; pr.phase = 0; pr.more = true;
    uint32_t head = 0, tail = 0;     // ring indices (warp-uniform registers mirror of W.ring_*)

    for (;;) {
        __syncwarp();
        // ------------------------------------------------------------------ produce when lanes would starve
        const unsigned idle_mask = __ballot_sync(FULL, !L.busy);
        const uint32_t n_idle = __popc(idle_mask), avail = tail - head;
        // lanes start new lines together (batch-synchronous refill): lines of one stream share their structure, so lanes
        // that start together stay in lockstep and share the divergent action code
        const bool all_idle = n_idle == 32;
        bool want = all_idle && (avail < 32u) && (pr.phase != 0 || pr.more) && (RING - avail >= (uint32_t)LT_MAX);
        if (want && pr.phase == 0) {
            // a free segment slot?
            int fs = -1;
            if (lane == 0) for (int k = 0; k < SEGSLOTS; k++) if (!W.slots[k].used) { fs = k; break; }
            fs = __shfl_sync(FULL, fs, 0);
            if (fs < 0) want = false;
            else {
                uint32_t s = 0;
                if (lane == 0) s = atomicAdd(&P.ctr->ticket, 1u);
                s = __shfl_sync(FULL, s, 0);
                if (s >= P.n_segs) { pr.more = false; want = false; }
                else {
                    pr.s = s; pr.seg = P.segs[s]; pr.mode = pr.seg.mode; if (pr.mode & SSE_MODE_R) pr.mode |= SSE_MODE_PARSE;
                    pr.conn = pr.seg.conn; pr.slot = (uint32_t)fs;
                    pr.rc.have_first = false; pr.rc.last_idx = SSE_NONE;
                    pr.rc.first.frame_first = pr.rc.first.frame_count = pr.rc.first.rec_first = pr.rc.first.rec_count = 0; pr.rc.first.next = SSE_NONE;
                    const ConnState cst = P.conns[pr.conn];
                    if (cst.flags & (CONN_FINISHED | CONN_DEAD)) {
                        if (lane == 0) {
                            sse_seg_result r; r.run = pr.rc.first; r.carry_len = cst.carry_len;
                            r.flags = (cst.flags & CONN_DEAD) ? SSE_SEG_DEAD : SSE_SEG_FINISHED; r.reserved = 0;
                            P.seg_results[s] = r;
                        }
                        continue;
                    }
                    if (lane == 0) {
                        SegSlot &sl = W.slots[fs];
                        sl.used = 1; sl.seg = s; sl.conn = pr.conn; sl.pending = 1; sl.term = ~0ull; sl.rmode = (pr.mode & SSE_MODE_R) ? 1u : 0u;
                    }
                    pr.in_long = (cst.flags & CONN_LONG) != 0;
                    pr.clen = (int)cst.carry_len; pr.pend = 0; pr.consumed = 0; pr.dead = pr.overflow = false;
                    if (!pr.in_long && pr.clen > 0) {
                        const int A = (pr.clen + 15) & ~15;
                        copy_g2s_bytes(buf + (A - pr.clen), P.carry + (size_t)pr.conn * P.carry_slot, pr.clen);
                        pr.pend = pr.clen; pr.clen = 0;
                    }
                    pr.phase = 1;
                    __syncwarp();
                }
            }
        }
        if (want && pr.phase != 0) {
            uint8_t *slot_mem = P.carry + (size_t)pr.conn * P.carry_slot;
            const int in_len = (int)pr.seg.in_len;
            if (pr.phase == 1) {           // ---- load the next window
                const int A = (pr.pend + 15) & ~15;
                pr.base = A - pr.pend;
                const int nload = min(in_len - pr.consumed, V2_BUF - A);
                {
                    const uint4 *g = reinterpret_cast<const uint4 *>(P.in + pr.seg.in_off + pr.consumed);
                    uint4 *d = reinterpret_cast<uint4 *>(buf + A);
                    const int nv = (nload + 15) >> 4;
                    for (int i = lane; i < nv; i += 32) d[i] = __ldg(g + i);
                }
                pr.consumed += nload;
                pr.fill = A + nload;
                pr.pos = pr.base;
                __syncwarp();
                pr.phase = 2;
                if (pr.in_long) {
                    int q = -1;
                    for (int i0 = pr.base; i0 < pr.fill && q < 0; i0 += 32) {
                        int i = i0 + (int)lane;
                        unsigned m = __ballot_sync(FULL, i < pr.fill && buf[i] == '\n');
                        if (m) q = i0 + __ffs(m) - 1;
                    }
                    const int take = (q < 0) ? (pr.fill - pr.base) : (q - pr.base + 1);
                    if (pr.clen + take > (int)P.carry_slot) { pr.dead = true; pr.phase = 3; }
                    else {
                        copy_s2g_bytes(slot_mem + pr.clen, buf + pr.base, take);
                        pr.clen += take;
                        __syncwarp();
                        if (q < 0) { pr.pend = 0; pr.phase = (pr.consumed >= in_len) ? 3 : 1; }
                        else {
                            if (!process_long_line(P, pr.rc, slot_mem, pr.clen, pr.mode)) { pr.overflow = true; pr.phase = 3; }
                            __syncwarp();
                            pr.in_long = false; pr.clen = 0; pr.pos = q + 1;
                        }
                    }
                }
            } else if (pr.phase == 2) {    // ---- one round of up to LT_MAX lines
                const uint32_t mode = pr.mode;
                const int pos = pr.pos, fill = pr.fill;
                if (lane == 0) W.done_cnt = 0;
                __syncwarp();
                int n_lines = 0;
                for (int g0 = pos & ~15; g0 < fill && n_lines < LT_MAX; g0 += 512) {
                    int off = g0 + (int)lane * 16;
                    uint32_t nlm = 0, brm = 0;
                    if (off < fill) {
                        uint4 v = *reinterpret_cast<const uint4 *>(buf + off);
                        nlm = eqmask16(v, 0x0A0A0A0Au);
                        if (mode & SSE_MODE_R) brm = eqmask16(v, 0x5B5B5B5Bu);
                        uint32_t valid = 0xFFFFu;
                        if (off < pos) valid &= 0xFFFFu << (pos - off);
                        if (off + 16 > fill) valid &= 0xFFFFu >> (off + 16 - fill);
                        nlm &= valid; brm &= valid;
                    }
                    while (brm) {
                        int bpos = off + __ffs(brm) - 1;
                        brm &= brm - 1;
                        if (bpos + 6 <= fill && is_done_at(buf + bpos)) {
                            uint32_t k = atomicAdd(&W.done_cnt, 1u);
                            if (k < DONE_MAX) W.done_pos[k] = (uint16_t)bpos;
                        }
                    }
                    uint32_t cnt = __popc(nlm), pre = cnt;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, pre, d); if ((int)lane >= d) pre += t; }
                    uint32_t total = __shfl_sync(FULL, pre, 31);
                    uint32_t idx = (uint32_t)n_lines + pre - cnt;
                    while (nlm) {
                        int bpos = off + __ffs(nlm) - 1;
                        nlm &= nlm - 1;
                        if (idx < LT_MAX) W.lt[idx].nl = (uint16_t)bpos;
                        idx++;
                    }
                    n_lines += (int)total;
                }
                __syncwarp();
                if (n_lines == 0) {
                    // ---- window exhausted: tail handling
                    const int tail_len = fill - pos;
                    if (pr.consumed >= in_len) {
                        if (tail_len > (int)P.carry_slot) pr.dead = true;
                        else { if (tail_len > 0) copy_s2g_bytes(slot_mem, buf + pos, tail_len); pr.clen = tail_len; }
                        pr.phase = 3;
                    } else if (tail_len >= V2_BUF - 32) {
                        if (tail_len > (int)P.carry_slot) { pr.dead = true; pr.phase = 3; }
                        else { copy_s2g_bytes(slot_mem, buf + pos, tail_len); pr.clen = tail_len; pr.in_long = true; pr.pend = 0; pr.phase = 1; }
                    } else {
                        // move the tail to the front so that it ends at a 16-byte aligned offset (pos > 32 > destination here)
                        const int A2 = (tail_len + 15) & ~15;
                        uint8_t *d = buf + (A2 - tail_len);
                        const uint8_t *sp = buf + pos;
                        if (d != sp) {
                            for (int i0 = 0; i0 < tail_len; i0 += 32) {
                                int i = i0 + (int)lane;
                                uint8_t v = (i < tail_len) ? sp[i] : (uint8_t)0;
                                __syncwarp();
                                if (i < tail_len) d[i] = v;
                                __syncwarp();
                            }
                        }
                        pr.pend = tail_len; pr.phase = 1;
                    }
                    __syncwarp();
                } else {
                    if (n_lines > LT_MAX) n_lines = LT_MAX;
                    const bool done_ovf = W.done_cnt > DONE_MAX;
                    const int n_done = min((int)W.done_cnt, DONE_MAX);
                    uint32_t my_flen[2] = { 0, 0 }, my_kind[2] = { 0, 0 }, my_parse[2] = { 0, 0 };
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        int i = (int)lane + 32 * h;
                        if (i < n_lines) {
                            int ls = (i == 0) ? pos : (int)W.lt[i - 1].nl + 1;
                            int nl = W.lt[i].nl;
                            int a = ls, b = nl + 1;
                            uint32_t kind, parse = 0; int src_s = ls, pay_s = ls, pay_e = ls, flen = 0;
                            if (mode & SSE_MODE_R) {
                                trim_space(buf, a, b);
                                bool has_done = false;
                                for (int k = 0; k < n_done; k++) { int d = W.done_pos[k]; if (d >= a && d + 6 <= b) has_done = true; }
                                if (done_ovf && !has_done) for (int k = a; k + 6 <= b; k++) if (is_done_at(buf + k)) { has_done = true; break; }
                                bool pref = is_data_prefix(buf + a, b - a);
                                if (has_done) {
                                    pay_s = pref ? a + 6 : a; pay_e = b;
                                    bool exact = (pay_e - pay_s) == 6 && is_done_at(buf + pay_s);
                                    kind = exact ? K_DONE_EXACT : K_DONE; parse = 1;
                                } else if (pref && b - a > 6) {
                                    kind = K_EMIT; src_s = a; pay_s = a + 6; pay_e = b; flen = (b - a) + 2; parse = 1;
                                } else kind = K_DROP;
                            } else {
                                kind = K_EMIT; src_s = ls; flen = nl + 1 - ls;
                                if ((mode & SSE_MODE_PARSE) && is_data_prefix(buf + ls, nl + 1 - ls)) { parse = 1; pay_s = ls + 6; pay_e = nl; }
                            }
                            LineEnt &e = W.lt[i];
                            e.src_s = (uint16_t)src_s; e.pay_s = (uint16_t)pay_s; e.pay_e = (uint16_t)pay_e;
                            e.flen = (uint16_t)flen; e.kind = (uint8_t)kind; e.parse = (uint8_t)parse;
                            my_flen[h] = (uint32_t)flen; my_kind[h] = kind; my_parse[h] = parse;
                        }
                    }
                    uint32_t pre_b[2], pre_f[2], pre_r[2], pre_q[2], tot_b = 0, tot_f = 0, tot_r = 0, tot_q = 0;
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint32_t vb = my_flen[h], vf = my_flen[h] ? 1u : 0u, vr = my_parse[h];
                        uint32_t vq = (my_parse[h] && my_kind[h] == K_EMIT) ? 1u : 0u;
                        uint32_t sb = vb, sf = vf | (vr << 8) | (vq << 16);   // three small counters in one scan
                        #pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            uint32_t tb = __shfl_up_sync(FULL, sb, d), tf = __shfl_up_sync(FULL, sf, d);
                            if ((int)lane >= d) { sb += tb; sf += tf; }
                        }
                        pre_b[h] = tot_b + sb - vb;
                        pre_f[h] = tot_f + (sf & 0xFF) - vf; pre_r[h] = tot_r + ((sf >> 8) & 0xFF) - vr; pre_q[h] = tot_q + ((sf >> 16) & 0xFF) - vq;
                        const uint32_t lb = __shfl_sync(FULL, sb, 31), lf = __shfl_sync(FULL, sf, 31);
                        tot_b += lb; tot_f += lf & 0xFF; tot_r += (lf >> 8) & 0xFF; tot_q += (lf >> 16) & 0xFF;
                    }
                    uint32_t ob = 0, fb = 0, rb = 0;
                    if (lane == 0) {
                        if (tot_b) ob = atomicAdd(&P.ctr->out_bytes, (tot_b + 15u) & ~15u);   // keep every round 16-byte aligned
                        if (tot_f) fb = atomicAdd(&P.ctr->n_frames, tot_f);
                        if (tot_r) rb = atomicAdd(&P.ctr->n_recs, tot_r);
                        if (tot_q) atomicAdd(&W.slots[pr.slot].pending, (int)tot_q);
                    }
                    ob = __shfl_sync(FULL, ob, 0); fb = __shfl_sync(FULL, fb, 0); rb = __shfl_sync(FULL, rb, 0);
                    if (ob + tot_b + 16 > P.cap_out || fb + tot_f > P.cap_frames || rb + tot_r > P.cap_recs || fb + tot_f >= (1u << 29)) {
                        if (lane == 0) { sse_overflow(P.ctr, SSE_OVF_OUT); if (tot_q) atomicSub(&W.slots[pr.slot].pending, (int)tot_q); }
                        pr.overflow = true; pr.phase = 3;
                    } else {
                        #pragma unroll
                        for (int h = 0; h < 2; h++)
                            if (my_flen[h]) { sse_frame f; f.off = ob + pre_b[h]; f.len = my_flen[h]; P.frames[fb + pre_f[h]] = f; }
                        __syncwarp();
                        {   // warp-cooperative serializer
                            uint32_t o = ob;
                            for (int i = 0; i < n_lines; i++) {
                                const LineEnt e = W.lt[i];
                                if (!e.flen) continue;
                                uint8_t *dst = P.out + o;
                                const uint8_t *sp = buf + e.src_s;
                                if (mode & SSE_MODE_R) {
                                    const int body = (int)e.flen - 2;     // "data: " + payload, contiguous in the window
                                    copy_s2g_vec(dst, sp, body);
                                    if (lane < 2) dst[body + lane] = (uint8_t)'\n';
                                } else copy_s2g_vec(dst, sp, (int)e.flen);
                                o += e.flen;
                            }
                        }
                        // work items for the consumer lanes; swallowed [DONE] lines are decoded here (rare)
                        #pragma unroll 1
                        for (int h = 0; h < 2; h++) {
                            int i = (int)lane + 32 * h;
                            if (i < n_lines && my_parse[h]) {
                                const LineEnt e = W.lt[i];
                                if (my_kind[h] == K_EMIT) {
                                    const uint32_t plen = (uint32_t)(e.pay_e - e.pay_s);
                                    uint4 it;
                                    it.x = ob + pre_b[h] + (uint32_t)(e.pay_s - e.src_s);
                                    it.y = plen;
                                    it.z = rb + pre_r[h];
                                    it.w = (fb + pre_f[h]) | (pr.slot << 29);
                                    W.ring[(tail + pre_q[h]) & (RING - 1)] = it;
                                } else {
                                    sse_rec r;
                                    r.frame = SSE_NONE; r.content_off = r.content_len = 0; r.tc_first = SSE_NONE; r.tc_count = 0; r.n_choices = 0;
                                    r.usage = SSE_NONE; r.payload_len = (uint32_t)(e.pay_e - e.pay_s);
                                    if (my_kind[h] == K_DONE_EXACT) r.flags = SSE_F_DONE_LINE | SSE_F_DONE_EXACT;
                                    else {
                                        ParseCtx cx; cx.sm = buf; cx.P = &P; cx.S = &c_schema; cx.emitted = false; cx.out_delta = 0;
                                        ParseOut po;
                                        decode_chunk(cx, e.pay_s, e.pay_e, po);
                                        r.flags = po.flags | SSE_F_DONE_LINE;
                                        r.content_off = po.content_off; r.content_len = po.content_len;
                                        r.tc_first = po.tc_first; r.tc_count = (uint16_t)min(po.tc_count, 0xFFFFu);
                                        r.n_choices = (uint16_t)min(po.n_choices, 0xFFFFu); r.usage = po.usage;
                                    }
                                    P.recs[rb + pre_r[h]] = r;
                                }
                            }
                        }
                        __threadfence_block();
                        __syncwarp();
                        tail += tot_q;
                        append_run(P, pr.rc, fb, tot_f, rb, tot_r);
                        pr.pos = (int)W.lt[n_lines - 1].nl + 1;
                    }
                }
            }
            if (pr.phase == 3) {           // ---- segment finished: connection state + draft result, release the slot's bias
                if (lane == 0) {
                    ConnState ns;
                    ns.carry_len = pr.dead ? 0u : (uint32_t)pr.clen;
                    ns.flags = (pr.dead ? CONN_DEAD : 0u) | ((pr.in_long && !pr.dead) ? CONN_LONG : 0u);
                    P.conns[pr.conn] = ns;
                    sse_seg_result r; r.run = pr.rc.first; r.carry_len = ns.carry_len;
                    r.flags = pr.dead ? (SSE_SEG_LINE_TOO_LONG | SSE_SEG_DEAD) : 0u; r.reserved = 0;
                    P.seg_results[pr.s] = r;
                    __threadfence_block();
                    SegSlot &sl = W.slots[pr.slot];
                    if (atomicSub(&sl.pending, 1) == 1) v2_finalize_segment(P, sl);
                }
                __syncwarp();
                pr.phase = 0;
            }
            continue;
        }
        // ------------------------------------------------------------------ nothing to produce: finished?
        if (avail == 0 && n_idle == 32) {
            if (pr.phase == 0 && !pr.more) break;
            // all lanes idle, ring empty, but the producer could not run (ring space / slots): cannot happen with an
            // empty ring, except when every segment slot is still marked used by a finalize in flight
            continue;
        }
        // ------------------------------------------------------------------ idle lanes take work items
        if (all_idle) {
            const uint32_t rank = lane;
            if (rank < avail) {
                const uint4 it = W.ring[(head + rank) & (RING - 1)];
                L.p = it.x; L.pe = it.x + it.y; L.rec = it.z; L.frame = it.w & 0x1FFFFFFFu;
                L.slot = it.w >> 29; L.plen = it.y;
                L.st = S_VAL; L.depth = L.skip = L.sd = 0; L.cur = TY_ROOT | (N_ROOT << 4); L.km = TRIE_ROOT; L.slen = 0;
                L.sf = W.slots[it.w >> 29].rmode ? SF_RMODE : 0u;
                L.choices_count = L.n_choices = 0; L.finish = SSE_FIN_NONE; L.ct = L.ct1 = L.sstk = 0;
                L.content_off = L.content_len = 0; L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE;
                S.u_prompt = S.u_completion = S.u_total = 0;
                L.busy = true;
                if (L.p < L.pe) L.win = ldwin16<false>(P.out, L.p);
            }
        }
        if (all_idle) head += min(avail, 32u);
        // ------------------------------------------------------------------ automaton: rounds of K plain steps, then
        // every lane that stopped at an action runs it (all lanes dispatch together: the divergent part is shared)
        #pragma unroll 1
        for (int round = 0; round < ROUNDS; round++) {
            v2_round<false>(P, T, L, S, nullptr);
            if (L.busy && L.p >= L.pe) {
                SegSlot &sl = W.slots[L.slot];
                if (v2_finish_line(P, L, S, nullptr)) atomicMin(&sl.term, ((unsigned long long)L.rec << 32) | L.frame);
                __threadfence_block();
                if (atomicSub(&sl.pending, 1) == 1) v2_finalize_segment(P, sl);
                L.p = L.pe = 0;
            }
        }
    }
}


// ---------------------------------------------------------------- split pipeline, stage 2: decode
// Persistent warps pull 32 work items at a time (lines of the same stream are adjacent, so the lanes of a batch walk
// near-identical structure in lockstep); no producer code and no line window in this kernel: small instruction
// footprint, shared memory only for the tables and the cold per-lane state.
#ifndef SSE_V3_WARPS
#define SSE_V3_WARPS 32
#endif
constexpr int V3_WARPS = SSE_V3_WARPS;

struct CtaSmem3 {
    DfaTables T;
    LaneScratch ls[V3_WARPS * 32];
    LaneJobs jobs[V3_WARPS * 32];
};
static_assert(sizeof(CtaSmem3) <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(V3_WARPS * 32, 1)
sse_decode_kernel(const __grid_constant__ KParams P, const DfaTables *__restrict__ gT) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaSmem3 &cs = *reinterpret_cast<CtaSmem3 *>(smem_raw);
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(gT);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&cs.T);
        for (int i = threadIdx.x; i < (int)(sizeof(DfaTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const DfaTables &T = cs.T;
    LaneScratch &S = cs.ls[threadIdx.x];
    LaneJobs *J = &cs.jobs[threadIdx.x];
    LaneJobs *Jw = &cs.jobs[threadIdx.x & ~31u];   // this warp's 32 queues
    J->n = 0;
    const uint32_t lane = lane_id();
    const uint32_t n_items = min(P.ctr->n_items, P.cap_items);

    Lane L; L.busy = false; L.p = L.pe = 0; L.win = make_uint4(0, 0, 0, 0);
    L.st = S_END; L.depth = L.skip = L.sd = 0; L.cur = 0; L.km = TRIE_ROOT; L.slen = 0; L.sf = 0; L.choices_count = L.n_choices = 0;
    L.finish = 0; L.ct = L.ct1 = L.sstk = 0; L.content_off = L.content_len = 0; L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE;
    L.rec = L.frame = L.slot = L.plen = 0;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&P.ctr->item_ticket, 32u);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n_items) break;
        const uint32_t idx = base + lane;
        if (idx < n_items) {
            const uint4 it = P.items_sorted[idx];
            L.p = it.x; L.plen = it.y & 0x00FFFFFFu; L.pe = it.x + L.plen; L.rec = it.z; L.slot = it.w;   // slot: segment index
            L.frame = P.recs[it.z].frame;
           
            L.st = S_VAL; L.depth = L.skip = L.sd = 0; L.cur = TY_ROOT | (N_ROOT << 4); L.km = TRIE_ROOT; L.slen = 0;
            L.sf = ((it.y & 0x80000000u) ? SF_RMODE : 0u) | ((it.y & 0x40000000u) ? SF_DONELINE : 0u);
            L.choices_count = L.n_choices = 0; L.finish = SSE_FIN_NONE; L.ct = L.ct1 = L.sstk = 0;
            L.content_off = L.content_len = 0; L.tc_count = 0; L.tc_first = L.tc_prev = SSE_NONE;
            S.u_prompt = S.u_completion = S.u_total = 0;
            L.busy = true;
            if (L.p < L.pe) L.win = ldwin16<true>(P.out, L.p);
        }
        while (__any_sync(FULL, L.busy)) {
            #pragma unroll 1
            for (int round = 0; round < ROUNDS; round++) {
                v2_round<true>(P, T, L, S, J);
                if (L.busy && L.p >= L.pe) {
                    if (v2_finish_line(P, L, S, J)) atomicMin(&P.seg_term[L.slot], L.rec);   // agent.go:235-242, resolved in stage 3
                    L.p = L.pe = 0;
                }
            }
            // strings that need unquoting were queued by the lanes: decode them with the whole warp
            __syncwarp();
            unsigned jm = __ballot_sync(FULL, J->n > 0);
            while (jm) {
                const int leader = __ffs(jm) - 1;
                jm &= jm - 1;
                LaneJobs &LJ = Jw[leader];
                const uint32_t nj = LJ.n;
                for (uint32_t k = 0; k < nj; k++) {
                    const UnquoteJob jb = LJ.j[k];
                    warp_unquote(P.out, jb.s, jb.e, P.text + jb.dst, jb.patch);
                }
                __syncwarp();
                if ((int)lane == leader) LJ.n = 0;
            }
            __syncwarp();
        }
    }
}


// ---------------------------------------------------------------- split pipeline, stage 1b: order the work items
// A decode warp takes 32 items and runs until the longest of them is done, so a batch of mixed lengths idles most lanes
// (payloads are lognormal 96..4096 B: in arrival order a batch is 34 % busy). Counting sort by payload length / 64, longest
// bucket first, and within a length bucket by shape class (provider hint x position of the line in its round): batches are then
// 94 % busy, their lanes stop at the same actions, and the long lines do not land on the tail of the kernel.
constexpr int N_BUCKETS = SSE_N_BUCKETS;
__device__ __forceinline__ uint32_t item_bucket(const KParams &P, uint32_t y) {
    const uint32_t cls = (y >> 24) & 31u;
    return (((uint32_t)SSE_LEN_BUCKETS - 1u - min((y & 0x00FFFFFFu) >> SSE_LEN_SHIFT, (uint32_t)SSE_LEN_BUCKETS - 1u)) << 5) | cls;
}
__global__ void sse_bucket_hist_kernel(const KParams P) {
    __shared__ uint32_t h[N_BUCKETS];
    for (int b = threadIdx.x; b < N_BUCKETS; b += blockDim.x) h[b] = 0;
    __syncthreads();
    const uint32_t n = min(P.ctr->n_items, P.cap_items);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&h[item_bucket(P, P.items[i].y)], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < N_BUCKETS; b += blockDim.x) if (h[b]) atomicAdd(&P.ctr->class_count[b], h[b]);
}
constexpr int SCAN_TPB = 1024, SCAN_PER = N_BUCKETS / SCAN_TPB;
static_assert(N_BUCKETS % SCAN_TPB == 0, "bucket count");
__global__ void __launch_bounds__(SCAN_TPB) sse_bucket_scan_kernel(const KParams P) {
    __shared__ uint32_t wsum[32];
    const uint32_t t = threadIdx.x, lane = t & 31u;
    uint32_t c[SCAN_PER], mine = 0;
    #pragma unroll
    for (int k = 0; k < SCAN_PER; k++) { c[k] = P.ctr->class_count[t * SCAN_PER + k]; mine += c[k]; }
    uint32_t incl = mine;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if ((int)lane >= d) incl += v; }
    if (lane == 31) wsum[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        const uint32_t w = wsum[t];
        uint32_t wi = w;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, wi, d); if ((int)lane >= d) wi += v; }
        wsum[t] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[t >> 5] + incl - mine;
    #pragma unroll
    for (int k = 0; k < SCAN_PER; k++) { P.ctr->class_cursor[t * SCAN_PER + k] = run; run += c[k]; }
}
constexpr int SCATTER_TPB = 256, SCATTER_IPT = 8;   // one block places 2048 consecutive items
__global__ void __launch_bounds__(SCATTER_TPB) sse_bucket_scatter_kernel(const KParams P) {
    __shared__ uint32_t h[N_BUCKETS], base[N_BUCKETS];
    const uint32_t n = min(P.ctr->n_items, P.cap_items);
    for (uint32_t blk = blockIdx.x * (SCATTER_TPB * SCATTER_IPT); blk < n; blk += gridDim.x * (SCATTER_TPB * SCATTER_IPT)) {
        for (int b = threadIdx.x; b < N_BUCKETS; b += SCATTER_TPB) h[b] = 0;
        __syncthreads();
        uint4 it[SCATTER_IPT]; uint32_t rank[SCATTER_IPT];
        #pragma unroll
        for (int k = 0; k < SCATTER_IPT; k++) {
            const uint32_t i = blk + k * SCATTER_TPB + threadIdx.x;
            if (i < n) { it[k] = P.items[i]; rank[k] = atomicAdd(&h[item_bucket(P, it[k].y)], 1u); }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < N_BUCKETS; b += SCATTER_TPB) if (h[b]) base[b] = atomicAdd(&P.ctr->class_cursor[b], h[b]);
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < SCATTER_IPT; k++) {
            const uint32_t i = blk + k * SCATTER_TPB + threadIdx.x;
            if (i < n) {
                const uint32_t pos = base[item_bucket(P, it[k].y)] + rank[k];
                P.items_sorted[pos] = it[k];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- split pipeline, stage 3: early termination
// One thread per segment: cut the segment's runs after the terminating chunk and mark the connection finished
// (everything after it is never read by the reference, mcp/agent.go:235-242 and :169).
__global__ void sse_finalize_kernel(const KParams P) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_segs) return;
    const uint32_t trec = P.seg_term[s];
    if (trec == SSE_NONE) return;
    const uint32_t tframe = P.recs[trec].frame;
    sse_seg_result r = P.seg_results[s];
    sse_run *run = &r.run;
    for (;;) {
        if (trec >= run->rec_first && trec < run->rec_first + run->rec_count) {
            run->rec_count = trec - run->rec_first + 1;
            run->frame_count = tframe - run->frame_first + 1;
            run->next = SSE_NONE;
            break;
        }
        if (run->next == SSE_NONE) break;
        run = &P.runs[run->next];
    }
    r.flags |= SSE_SEG_TERMINATED;
    r.carry_len = 0;
    P.seg_results[s] = r;
    ConnState ns; ns.carry_len = 0; ns.flags = CONN_FINISHED;
    P.conns[P.segs[s].conn] = ns;
}

DfaTables *g_tables_dev[16] = { nullptr };

} // namespace

int sse_v2_prepare(int device) {
    if (device < 0 || device >= 16) return (int)cudaErrorInvalidValue;
    if (g_tables_dev[device]) return 0;
    static Schema keep;   // field names must outlive build_tables
    cudaError_t e = cudaMemcpyFromSymbol(&keep, c_schema, sizeof keep);
    if (e != cudaSuccess) return (int)e;
    static ssetab::FieldSrc fs[N_FIELDS];
    int n = 0;
    for (int node = 0; node < N_COUNT; node++)
        for (int k = 0; k < keep.cnt[node]; k++) {
            const FieldDef &f = keep.f[keep.first[node] + k];
            fs[n].node = (uint8_t)node; fs[n].name = f.name; fs[n].ty = f.ty; fs[n].sub = f.sub; fs[n].tgt = f.tgt;
            n++;
        }
    static const char *fin_names[] = { "stop", "tool_calls", "length", "content_filter", "function_call" };
    static const uint8_t fin_vals[] = { SSE_FIN_STOP, SSE_FIN_TOOL_CALLS, SSE_FIN_LENGTH, SSE_FIN_CONTENT_FILTER, SSE_FIN_FUNCTION_CALL };
    static ssetab::DfaTables T;
    if (ssetab::build_tables(T, fs, n, fin_names, fin_vals, 5) != 0) return (int)cudaErrorInvalidValue;
    DfaTables *d = nullptr;
    e = cudaMalloc((void **)&d, sizeof T);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d, &T, sizeof T, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(sse_stream_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem2));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(sse_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem3));
    if (e != cudaSuccess) return (int)e;
    g_tables_dev[device] = d;
    return 0;
}

int sse_launch_stream_kernel_v2(const KParams &p, void *stream, int sm_count, int device) {
    int grid = sm_count;   // persistent: one 16-warp CTA per SM
    int need = (int)((p.n_segs + V2_WARPS - 1) / V2_WARPS);
    if (need < grid) grid = need > 0 ? need : 1;
    sse_stream_kernel_v2<<<grid, V2_WARPS * 32, sizeof(CtaSmem2), (cudaStream_t)stream>>>(p, g_tables_dev[device]);
    return (int)cudaGetLastError();
}

int sse_launch_decode_finalize(const KParams &p, void *stream, int sm_count, int device) {
    sse_bucket_hist_kernel<<<sm_count * 2, 512, 0, (cudaStream_t)stream>>>(p);
    sse_bucket_scan_kernel<<<1, SCAN_TPB, 0, (cudaStream_t)stream>>>(p);
    sse_bucket_scatter_kernel<<<sm_count * 2, SCATTER_TPB, 0, (cudaStream_t)stream>>>(p);
    sse_decode_kernel<<<sm_count, V3_WARPS * 32, sizeof(CtaSmem3), (cudaStream_t)stream>>>(p, g_tables_dev[device]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const int tpb = 256;
    sse_finalize_kernel<<<(p.n_segs + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
