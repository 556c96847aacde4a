// sse_kernel.cu -- stage 1 of the default pipeline (produce).
//
// One persistent warp per connection segment (dynamic ticket):
//   stage    carry tail (HBM, per connection) + new segment bytes -> warp-private shared-memory window
//   split    SWAR/ballot newline scan -> line table            (provider.go:322 ReadBytes('\n'))
//   classify strings.TrimSpace / Contains "[DONE]" / HasPrefix "data: "   (agent.go:178-193), or verbatim (mode P)
//   emit     frame table; a frame that stands in the input arena as it must be sent is a span of it (zero-copy), the others
//            go through the warp-cooperative serializer into the out arena      (agent.go:195 / routes.go:613)
//   items    a stub record + a 16-byte work item per line to decode (lines longer than the window are assembled in the carry
//            slot and become work items too); sse_kernel2.cu sorts the items, decodes them (sse_decode_kernel) and resolves
//            early termination (sse_finalize_kernel). The template parameter SPLIT is always true: the first-generation
//            kernel that decoded in place (SPLIT = false) was removed in round 2.
//   finish   carry update, per-segment result
//
// Pure integer/byte work, no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include "sse_common.cuh"

namespace {

// ---------------------------------------------------------------- the kernel
// stage 1 of the pipeline (no JSON decoder in this kernel)
#ifndef SSE_V1_MINB
#define SSE_V1_MINB 2
#endif
template <bool SPLIT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, SSE_V1_MINB)
sse_stream_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CtaSmem &cs = *reinterpret_cast<CtaSmem *>(smem_raw);
    {   // schema table: constant -> shared (lanes index it divergently)
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&c_schema);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&cs.schema);
        for (int i = threadIdx.x; i < (int)(sizeof(Schema) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    WarpSmem &W = cs.w[threadIdx.x >> 5];
    uint8_t *buf = W.buf;
    const uint32_t lane = lane_id();

    // tickets are taken one segment ahead: while a segment is processed, the next one's first window is on its way to L2
    uint32_t s_next = 0;
    if (lane == 0) s_next = atomicAdd(&P.ctr->ticket, 1u);
    s_next = __shfl_sync(FULL, s_next, 0);
    for (;;) {
        const uint32_t s = s_next;
        if (s >= P.n_segs) break;
        if (lane == 0) s_next = atomicAdd(&P.ctr->ticket, 1u);
        s_next = __shfl_sync(FULL, s_next, 0);
        if (s_next < P.n_segs) {
            const sse_seg nx = P.segs[s_next];
            const uint32_t nb = min(nx.in_len, (uint32_t)BUF);
            for (uint32_t o = lane * 128u; o < nb; o += 32u * 128u) prefetch_l2(P.in + nx.in_off + o);
        }

        const sse_seg seg = P.segs[s];
        uint32_t mode = seg.mode;
        if (mode & SSE_MODE_R) mode |= SSE_MODE_PARSE;
        const uint32_t conn = seg.conn;
        ConnState cst = P.conns[conn];
        uint8_t *slot = P.carry + (size_t)conn * P.carry_slot;
        RunChain rc; rc.have_first = false; rc.last_idx = SSE_NONE;
        rc.first.frame_first = rc.first.frame_count = rc.first.rec_first = rc.first.rec_count = 0; rc.first.next = SSE_NONE;
        uint32_t seg_flags = 0;
        if (SPLIT && lane == 0) P.seg_term[s] = SSE_NONE;

        if (cst.flags & (CONN_FINISHED | CONN_DEAD)) {
            if (lane == 0) {
                sse_seg_result r; r.run = rc.first; r.carry_len = cst.carry_len;
                r.flags = (cst.flags & CONN_DEAD) ? SSE_SEG_DEAD : SSE_SEG_FINISHED; r.reserved = 0;
                P.seg_results[s] = r;
            }
            continue;
        }

        bool in_long = (cst.flags & CONN_LONG) != 0;
        int clen = (int)cst.carry_len;     // bytes in the carry slot
        int pend = 0;                      // bytes at the window front (end at a 16-aligned offset)
        bool carry_front = false;          // the window front holds bytes of an earlier batch (not in this input arena)
        if (!in_long && clen > 0) {
            int A = (clen + 15) & ~15;
            copy_g2s_bytes(buf + (A - clen), slot, clen);
            pend = clen; clen = 0; carry_front = true;
        }
        const bool zero_copy = !(P.flags & SSE_FLAG_COPY_OUT);
        int consumed = 0;
        const int in_len = (int)seg.in_len;
        const uint8_t *src = P.in + seg.in_off;
        bool terminated = false, dead = false, overflow = false;

        for (;;) {   // windows
            const int A = (pend + 15) & ~15;
            const int base = A - pend;
            int room = BUF - A;
            int nload = min(in_len - consumed, room);
            {   // 16-byte vector copy HBM -> shared (segment offsets and A are 16-byte aligned)
                const uint4 *g = reinterpret_cast<const uint4 *>(src + consumed);
                uint4 *d = reinterpret_cast<uint4 *>(buf + A);
                int nv = (nload + 15) >> 4;
                for (int i = lane; i < nv; i += 32) d[i] = __ldg(g + i);
            }
            consumed += nload;
            const int fill = A + nload;
            // window position w >= zc_lo holds input byte seg.in_off + ... = w + in_delta (a tail moved to the front is input too)
            const int in_delta = (int)seg.in_off + (consumed - nload) - A;
            const int zc_lo = carry_front ? A : base;
            carry_front = false;
            __syncwarp();
            int pos = base;

            if (in_long) {
                // still inside a line longer than the window: look for its end only
                int q = -1;
                for (int i0 = base; i0 < fill && q < 0; i0 += 32) {
                    int i = i0 + (int)lane;
                    unsigned m = __ballot_sync(FULL, i < fill && buf[i] == '\n');
                    if (m) q = i0 + __ffs(m) - 1;
                }
                int take = (q < 0) ? (fill - base) : (q - base + 1);
                if (clen + take > (int)P.carry_slot) { dead = true; break; }
                copy_s2g_bytes(slot + clen, buf + base, take);
                clen += take;
                __syncwarp();
                if (q < 0) {
                    pend = 0;
                    if (consumed >= in_len) break;
                    continue;
                }
                if (!process_long_line(P, rc, slot, clen, mode, SPLIT ? (int)s : -1)) { overflow = true; break; }
                __syncwarp();
                in_long = false; clen = 0; pos = q + 1;
            }

            // ---- rounds of up to LT_MAX lines
            for (;;) {
                if (lane == 0) W.done_cnt = 0;
                __syncwarp();
                // split: newline scan, 16 bytes per lane per step
                int n_lines = 0;
                for (int g0 = pos & ~15; g0 < fill && n_lines < LT_MAX; g0 += 512) {
                    int off = g0 + (int)lane * 16;
                    uint32_t nlm = 0, brm = 0;
                    if (off < fill) {
                        uint4 v = *reinterpret_cast<const uint4 *>(buf + off);
                        nlm = eqmask16(v, 0x0A0A0A0Au);
                        if (mode & SSE_MODE_R) brm = done_candidates16(v);
                        // mask bytes outside [pos, fill)
                        uint32_t valid = 0xFFFFu;
                        if (off < pos) valid &= 0xFFFFu << (pos - off);
                        if (off + 16 > fill) valid &= 0xFFFFu >> (off + 16 - fill);
                        nlm &= valid; brm &= valid;
                    }
                    while (brm) {   // "[DONE]" candidates (agent.go:181)
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
                if (n_lines == 0) break;
                if (n_lines > LT_MAX) n_lines = LT_MAX;
                const bool done_ovf = W.done_cnt > DONE_MAX;
                const int n_done = min((int)W.done_cnt, DONE_MAX);

                // classify: one lane per line (2 passes for 64 lines)
                uint32_t my_flen[2] = { 0, 0 }, my_kind[2] = { 0, 0 }, my_parse[2] = { 0, 0 }, my_zc[2] = { 0, 0 }, my_clen[2] = { 0, 0 };
                #pragma unroll
                for (int h = 0; h < 2; h++) {
                    int i = (int)lane + 32 * h;
                    if (i < n_lines) {
                        int ls = (i == 0) ? pos : (int)W.lt[i - 1].nl + 1;
                        int nl = W.lt[i].nl;
                        int a = ls, b = nl + 1;
                        uint32_t kind, parse = 0, zc = 0; int src_s = ls, pay_s = ls, pay_e = ls, flen = 0;
                        if (mode & SSE_MODE_R) {
                            trim_space(buf, a, b);                          // agent.go:178-179
                            bool has_done = false;                          // agent.go:181
                            for (int k = 0; k < n_done; k++) { int d = W.done_pos[k]; if (d >= a && d + 6 <= b) has_done = true; }
                            if (done_ovf && !has_done) for (int k = a; k + 6 <= b; k++) if (is_done_at(buf + k)) { has_done = true; break; }
                            bool pref = is_data_prefix(buf + a, b - a);    // agent.go:186
                            if (has_done) {
                                // swallowed; still decoded for parseStreamingToolCalls (agent.go:182, :377-402)
                                pay_s = pref ? a + 6 : a; pay_e = b;
                                bool exact = pref && (pay_e - pay_s) == 6 && is_done_at(buf + pay_s);   // only `data: [DONE]` ends A6's loop (agent.go:385-396)
                                kind = exact ? K_DONE_EXACT : K_DONE; parse = 1;
                                zc = (zero_copy && pay_s >= zc_lo) ? 1u : 0u;   // split: decoded in place by the decode kernel
                            } else if (pref && b - a > 6) {                  // agent.go:190-197
                                kind = K_EMIT; src_s = a; pay_s = a + 6; pay_e = b; flen = (b - a) + 2; parse = 1;
                                // "data: " + payload + "\n\n" is already what the input holds when nothing was trimmed
                                // at the end and the separator line is empty (the usual upstream framing)
                                zc = (zero_copy && b == nl && a >= zc_lo && nl + 1 < fill && buf[nl + 1] == '\n') ? 1u : 0u;
                            } else kind = K_DROP;
                        } else {
                            kind = K_EMIT; src_s = ls; flen = nl + 1 - ls;   // routes.go:613 verbatim
                            zc = (zero_copy && ls >= zc_lo) ? 1u : 0u;
                            if ((mode & SSE_MODE_PARSE) && is_data_prefix(buf + ls, nl + 1 - ls)) { parse = 1; pay_s = ls + 6; pay_e = nl; }
                        }
                        LineEnt &e = W.lt[i];
                        e.src_s = (uint16_t)src_s; e.pay_s = (uint16_t)pay_s; e.pay_e = (uint16_t)pay_e;
                        // bytes to materialise: a frame that is not a span of the input; in the split pipeline also the payload of a
                        // swallowed "[DONE]"-containing line (agent.go:182: still parsed), which the decode kernel reads from an arena
                        const int clen = zc ? 0 : (flen ? flen : ((SPLIT && kind == K_DONE) ? pay_e - pay_s : 0));
                        e.flen = (uint16_t)flen; e.kind = (uint8_t)kind; e.parse = (uint8_t)parse; e.zc = (uint16_t)zc; e.clen = (uint16_t)clen;
                        my_flen[h] = (uint32_t)flen; my_kind[h] = kind; my_parse[h] = parse; my_zc[h] = zc; my_clen[h] = (uint32_t)clen;
                    }
                }
                // allocate: exclusive prefix (line order) of frame bytes / frame count / rec count
                uint32_t pre_b[2], pre_f[2], pre_r[2], tot_b = 0, tot_f = 0, tot_r = 0;
                #pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t vb = my_clen[h], vf = my_flen[h] ? 1u : 0u, vr = my_parse[h];
                    uint32_t sb = vb, sf = vf, sr = vr;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        uint32_t tb = __shfl_up_sync(FULL, sb, d), tf = __shfl_up_sync(FULL, sf, d), tr = __shfl_up_sync(FULL, sr, d);
                        if ((int)lane >= d) { sb += tb; sf += tf; sr += tr; }
                    }
                    pre_b[h] = tot_b + sb - vb; pre_f[h] = tot_f + sf - vf; pre_r[h] = tot_r + sr - vr;
                    tot_b += __shfl_sync(FULL, sb, 31); tot_f += __shfl_sync(FULL, sf, 31); tot_r += __shfl_sync(FULL, sr, 31);
                }
                // split pipeline: every line to decode is a work item; index = rank among this round's items
                uint32_t tot_q = 0, qb = 0;
                if (SPLIT) {
                    uint32_t my_q = 0;
                    #pragma unroll
                    for (int h = 0; h < 2; h++) my_q += (my_parse[h] && (my_kind[h] == K_EMIT || my_kind[h] == K_DONE)) ? 1u : 0u;
                    uint32_t pre_q = my_q;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, pre_q, d); if ((int)lane >= d) pre_q += t; }
                    tot_q = __shfl_sync(FULL, pre_q, 31);
                    pre_q -= my_q;
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        int i = (int)lane + 32 * h;
                        if (i < n_lines && my_parse[h] && (my_kind[h] == K_EMIT || my_kind[h] == K_DONE))
                            W.lt[i].rel = (uint16_t)(pre_q + ((h == 1 && my_parse[0] && (my_kind[0] == K_EMIT || my_kind[0] == K_DONE)) ? 1u : 0u));
                    }
                    __syncwarp();
                }
                uint32_t ob = 0, fb = 0, rb = 0;
                if (lane == 0) {     // the round's allocation: two 64-bit atomics (Counters: out_bytes | n_items, n_frames | n_recs)
                    const unsigned long long a0 = (unsigned long long)((tot_b + 15u) & ~15u) | ((unsigned long long)tot_q << 32);
                    const unsigned long long a1 = (unsigned long long)tot_f | ((unsigned long long)tot_r << 32);
                    unsigned long long r0 = 0, r1 = 0;
                    if (a0) r0 = atomicAdd(reinterpret_cast<unsigned long long *>(&P.ctr->out_bytes), a0);
                    if (a1) r1 = atomicAdd(reinterpret_cast<unsigned long long *>(&P.ctr->n_frames), a1);
                    ob = (uint32_t)r0; qb = (uint32_t)(r0 >> 32); fb = (uint32_t)r1; rb = (uint32_t)(r1 >> 32);
                }
                ob = __shfl_sync(FULL, ob, 0); fb = __shfl_sync(FULL, fb, 0); rb = __shfl_sync(FULL, rb, 0); qb = __shfl_sync(FULL, qb, 0);
                if (ob + tot_b + 16 > P.cap_out || fb + tot_f > P.cap_frames || rb + tot_r > P.cap_recs ||
                    (SPLIT && qb + tot_q > P.cap_items)) {
                    if (lane == 0) sse_overflow(P.ctr, SSE_OVF_OUT);
                    overflow = true; break;
                }
                // frame table
                #pragma unroll
                for (int h = 0; h < 2; h++)
                    if (my_flen[h]) {
                        sse_frame f; f.len = my_flen[h];
                        f.off = my_zc[h] ? P.in_base + (uint32_t)(in_delta + (int)W.lt[(int)lane + 32 * h].src_s) : ob + pre_b[h];
                        P.frames[fb + pre_f[h]] = f;
                    }
                __syncwarp();
                // emit: warp-cooperative serializer, frames back to back in line order (nothing to do when every frame of the
                // round is a span of the input)
                if (tot_b) {
                    uint32_t o = ob;
                    for (int i = 0; i < n_lines; i++) {
                        const LineEnt e = W.lt[i];
                        if (!e.clen) continue;
                        uint8_t *dst = P.out + o;
                        if (!e.flen) copy_s2g_vec(dst, buf + e.pay_s, (int)e.clen);   // payload of a swallowed line (no frame)
                        else if (mode & SSE_MODE_R) {
                            const int body = (int)e.flen - 2;      // "data: " + payload is contiguous in the window
                            copy_s2g_vec(dst, buf + e.src_s, body);
                            if (lane < 2) dst[body + lane] = (uint8_t)'\n';
                        } else copy_s2g_vec(dst, buf + e.src_s, (int)e.flen);
                        o += e.clen;
                    }
                }
                // decode: one lane per line
                uint32_t term_line = 0xFFFFu;
                constexpr int UNROLL_DEC = SPLIT ? 2 : 1;   // the produce stage only writes a stub and a work item here
                #pragma unroll UNROLL_DEC
                for (int h = 0; h < 2; h++) {
                    int i = (int)lane + 32 * h;
                    if (i < n_lines && my_parse[h]) {
                        const LineEnt e = W.lt[i];
                        ParseCtx cx; cx.sm = buf; cx.P = &P; cx.S = &cs.schema;
                        cx.emitted = my_kind[h] == K_EMIT;
                        cx.out_delta = e.zc ? (int64_t)P.in_base + in_delta
                                            : (int64_t)(ob + pre_b[h]) - (int64_t)(e.flen ? e.src_s : e.pay_s);
                        if (SPLIT && (my_kind[h] == K_EMIT || my_kind[h] == K_DONE)) {
                            const bool swallowed = my_kind[h] == K_DONE;
                            sse_rec stub;
                            stub.frame = swallowed ? SSE_NONE : fb + pre_f[h]; stub.flags = 0; stub.content_off = stub.content_len = 0; stub.tc_first = SSE_NONE;
                            stub.tc_count = 0; stub.n_choices = 0; stub.usage = SSE_NONE; stub.payload_len = (uint32_t)(e.pay_e - e.pay_s);
                            P.recs[rb + pre_r[h]] = stub;
                            uint4 it;
                            it.x = (uint32_t)(cx.out_delta + (int64_t)e.pay_s);
                            it.z = rb + pre_r[h];
                            // scheduling key of the decode kernel's item sort: provider hint x position of the line in its round; the
                            // second decoded line of a round (first content delta, the longest) sorts first: long work early
                            const uint32_t ord = min((uint32_t)e.rel, 7u);
                            const uint32_t cls = ((ord == 1u ? 0u : (ord == 0u ? 1u : ord)) << 2) | (seg.provider & 3u);
                            // bit 31: the line can terminate the stream (emitted, mode R); bit 30: swallowed line (SSE_F_DONE_LINE)
                            it.y = (uint32_t)(e.pay_e - e.pay_s) | (cls << 24) | (swallowed ? 0x40000000u : ((mode & SSE_MODE_R) ? 0x80000000u : 0u));
                            it.w = s;
                            P.items[qb + e.rel] = it;
                            continue;
                        }
                        ParseOut po;
                        // (split pipeline: only exact "[DONE]" payloads get here, the sequential decoder is not part of that kernel)
                        po.flags = 0; po.content_off = po.content_len = 0; po.tc_first = SSE_NONE; po.tc_count = po.n_choices = 0; po.usage = SSE_NONE;
                        sse_rec r;
                        r.frame = cx.emitted ? fb + pre_f[h] : SSE_NONE;
                        r.flags = po.flags;
                        if (my_kind[h] == K_DONE) r.flags |= SSE_F_DONE_LINE;
                        if (my_kind[h] == K_DONE_EXACT) r.flags = SSE_F_DONE_LINE | SSE_F_DONE_EXACT;
                        uint32_t fin = (po.flags & SSE_F_FINISH_MASK) >> SSE_F_FINISH_SHIFT;
                        if ((mode & SSE_MODE_R) && my_kind[h] == K_EMIT && (po.flags & SSE_F_JSON_OK) && po.n_choices > 0 &&
                            (fin == SSE_FIN_STOP || fin == SSE_FIN_TOOL_CALLS)) {            // agent.go:235-242
                            r.flags |= SSE_F_TERMINATES;
                            term_line = min(term_line, (uint32_t)i);
                        }
                        r.content_off = po.content_off; r.content_len = po.content_len;
                        r.tc_first = po.tc_first; r.tc_count = (uint16_t)min(po.tc_count, 0xFFFFu);
                        r.n_choices = (uint16_t)min(po.n_choices, 0xFFFFu);
                        r.usage = po.usage; r.payload_len = (uint32_t)(e.pay_e - e.pay_s);
                        if (my_kind[h] == K_DONE_EXACT) { r.content_off = r.content_len = 0; r.tc_first = SSE_NONE; r.tc_count = 0; r.n_choices = 0; r.usage = SSE_NONE; }
                        P.recs[rb + pre_r[h]] = r;
                    }
                }
                #pragma unroll
                for (int d = 16; d >= 1; d >>= 1) term_line = min(term_line, __shfl_xor_sync(FULL, term_line, d));
                uint32_t run_f = tot_f, run_r = tot_r;
                if (term_line != 0xFFFFu) {
                    // lines after the terminating chunk are never read by the reference: cut the run there
                    uint32_t cf = 0, cr = 0;
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        int i = (int)lane + 32 * h;
                        if (i < n_lines && (uint32_t)i <= term_line) { cf += my_flen[h] ? 1u : 0u; cr += my_parse[h]; }
                    }
                    #pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) { cf += __shfl_xor_sync(FULL, cf, d); cr += __shfl_xor_sync(FULL, cr, d); }
                    run_f = cf; run_r = cr; terminated = true;
                }
                append_run(P, rc, fb, run_f, rb, run_r);
                pos = (int)W.lt[n_lines - 1].nl + 1;
                __syncwarp();
                if (terminated) break;
            }
            if (terminated || overflow) break;

            const int tail = fill - pos;
            if (consumed >= in_len) {          // segment exhausted: hold the unterminated tail back
                if (tail > 0) {
                    if (tail > (int)P.carry_slot) { dead = true; break; }
                    copy_s2g_bytes(slot, buf + pos, tail);
                }
                clen = tail;
                break;
            }
            if (tail >= BUF - 32) {            // a whole window without '\n': assemble the line in HBM
                if (tail > (int)P.carry_slot) { dead = true; break; }
                copy_s2g_bytes(slot, buf + pos, tail);
                clen = tail; in_long = true; pend = 0;
                __syncwarp();
                continue;
            }
            // move the tail to the front so that it ends at a 16-byte aligned offset. The window was filled to
            // BUF (more segment bytes remain) and tail < BUF-32, so pos > 32 > destination: the move is toward
            // lower addresses and ascending 32-byte steps (read, sync, write) never clobber unread source bytes.
            {
                const int A2 = (tail + 15) & ~15;
                uint8_t *d = buf + (A2 - tail);
                const uint8_t *sp = buf + pos;
                if (d != sp) {
                    for (int i0 = 0; i0 < tail; i0 += 32) {
                        int i = i0 + (int)lane;
                        uint8_t v = (i < tail) ? sp[i] : (uint8_t)0;
                        __syncwarp();
                        if (i < tail) d[i] = v;
                        __syncwarp();
                    }
                }
                pend = tail;
            }
        }

        if (terminated) seg_flags |= SSE_SEG_TERMINATED;
        if (dead) seg_flags |= SSE_SEG_LINE_TOO_LONG | SSE_SEG_DEAD;
        if (lane == 0) {
            ConnState ns;
            ns.carry_len = (terminated || dead) ? 0u : (uint32_t)clen;
            ns.flags = (terminated ? CONN_FINISHED : 0u) | (dead ? CONN_DEAD : 0u) | ((in_long && !terminated && !dead) ? CONN_LONG : 0u);
            P.conns[conn] = ns;
            sse_seg_result r; r.run = rc.first; r.carry_len = ns.carry_len; r.flags = seg_flags; r.reserved = 0;
            P.seg_results[s] = r;
        }
        __syncwarp();
    }
}

} // namespace

template <bool SPLIT>
static int launch_v1(const KParams &p, void *stream, int sm_count) {
    static bool attr_set = false;
    static int ctas_per_sm = 0;
    const size_t smem = sizeof(CtaSmem);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sse_stream_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sse_stream_kernel<SPLIT>, WARPS_PER_CTA * 32, smem);
        if (e != cudaSuccess) return (int)e;
        ctas_per_sm = n > 0 ? n : 1;
        attr_set = true;
    }
    int grid = sm_count * ctas_per_sm;   // persistent: one resident wave, warps pull segments by ticket
    int need = (int)((p.n_segs + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    if (need < grid) grid = need > 0 ? need : 1;
    sse_stream_kernel<SPLIT><<<grid, WARPS_PER_CTA * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

int sse_launch_produce_kernel(const KParams &p, void *stream, int sm_count) { return launch_v1<true>(p, stream, sm_count); }
