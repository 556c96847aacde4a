/* orc_check.c -- full-size parity checker: per-connection digests of everything the path produces.
 * TEST INFRASTRUCTURE ONLY (tests/ and __graft_entry__.smoke()); never on the product path.
 *
 * Comparing 65,536 streams x ~11 events field by field from Python takes minutes, so both sides are reduced in C to the
 * same canonical byte string per connection and hashed (FNV-1a, 64 bit):
 *   frames digest : for every emitted frame, in order:  u32 len, bytes
 *   records digest: for every decoded line, in order:   u8 json_ok, u8 done, u8 done_exact, and when json_ok:
 *                   u32 min(n_choices, 65535), u32 finish, u8 has_usage [3 x i64], u32 content_len, content bytes,
 *                   u8 tool_calls_nonnil, u8 has_valid_tool_call, u32 min(tc_count, 65535), then per element
 *                   i64 index, u8 has_id [u32 len, bytes], u8 has_type [u32 len, bytes], u8 has_function,
 *                   u32 name_len, name, u32 args_len, args
 * orc_digest_streams runs the oracle (sse_oracle.c) over whole upstream bodies; orc_digest_result folds ONE batch result of
 * the GPU library (the plain C structs of include/sse_gpu.h) into the digests of the connections its segments belong to, so a
 * stream cut into any number of batches ends at the same digest as the oracle's single pass over it.
 * The field list is the one tests/util.py's check_stream compares (chunk_to_dict / rec_to_dict). */
#define _GNU_SOURCE
#include "sse_oracle.h"
#include "../include/sse_gpu.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t frames_h, recs_h;      /* running FNV-1a states (start: ORC_FNV_INIT) */
    uint64_t n_frames, n_recs, frame_bytes;
    uint32_t inexact;               /* GPU side: records flagged TOO_LONG / DEPTH_LIMIT (side-band not decoded) */
    uint32_t pad;
} orc_digest;

#define ORC_FNV_INIT 0xcbf29ce484222325ull
#define ORC_FNV_MUL 0x100000001b3ull

static inline uint64_t h_bytes(uint64_t h, const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= ORC_FNV_MUL; }
    return h;
}
static inline uint64_t h_u8(uint64_t h, uint32_t v) { uint8_t b = (uint8_t)v; return h_bytes(h, &b, 1); }
static inline uint64_t h_u32(uint64_t h, uint32_t v) { return h_bytes(h, &v, 4); }
static inline uint64_t h_i64(uint64_t h, int64_t v) { return h_bytes(h, &v, 8); }
static inline uint64_t h_str(uint64_t h, const uint8_t *p, uint32_t n) { h = h_u32(h, n); return h_bytes(h, p, n); }

void orc_digest_init(orc_digest *d, size_t n) {
    for (size_t i = 0; i < n; i++) { memset(&d[i], 0, sizeof d[i]); d[i].frames_h = d[i].recs_h = ORC_FNV_INIT; }
}

/* ---------------------------------------------------------------- oracle side */
static uint64_t h_chunk(uint64_t h, const orc_result *r, const orc_chunk *c, int done, int done_exact) {
    h = h_u8(h, c ? c->json_ok : 0); h = h_u8(h, (uint32_t)done); h = h_u8(h, (uint32_t)done_exact);
    if (!c || !c->json_ok) return h;
    h = h_u32(h, c->n_choices > 65535u ? 65535u : c->n_choices);
    h = h_u32(h, c->finish);
    h = h_u8(h, c->has_usage);
    if (c->has_usage) { h = h_i64(h, c->prompt); h = h_i64(h, c->completion); h = h_i64(h, c->total); }
    h = h_str(h, r->text + c->content.off, c->content.len);
    h = h_u8(h, c->tool_calls_nonnil); h = h_u8(h, c->has_valid_tool_call);
    h = h_u32(h, c->tc_count > 65535u ? 65535u : c->tc_count);
    for (uint32_t k = 0; k < c->tc_count && k < 65535u; k++) {
        const orc_tc *t = &r->tcs[c->tc_first + k];
        h = h_i64(h, t->index);
        h = h_u8(h, t->has_id);   if (t->has_id) h = h_str(h, r->text + t->id.off, t->id.len);
        h = h_u8(h, t->has_type); if (t->has_type) h = h_str(h, r->text + t->type.off, t->type.len);
        h = h_u8(h, t->has_function);
        h = h_str(h, r->text + t->name.off, t->name.len);
        h = h_str(h, r->text + t->args.off, t->args.len);
    }
    return h;
}

static void digest_one(orc_result *r, const uint8_t *body, size_t n, uint32_t mode, orc_digest *d, uint8_t *terminated, uint32_t *tail_len) {
    orc_result_clear(r);
    if (mode & 1u) orc_reframe_stream(r, body, n, 0);
    else orc_passthrough(r, body, n, (mode & 2u) != 0);
    for (size_t i = 0; i < r->n_lines; i++) {
        const orc_line *l = &r->lines[i];
        const orc_chunk *c = l->chunk != 0xFFFFFFFFu ? &r->chunks[l->chunk] : 0;
        if (mode & 1u) {
            if (l->kind == ORC_L_EMITTED) {
                d->frames_h = h_str(d->frames_h, r->out + l->out_off, l->out_len); d->n_frames++; d->frame_bytes += l->out_len;
                d->recs_h = h_chunk(d->recs_h, r, c, 0, 0); d->n_recs++;
            } else if (l->kind == ORC_L_DONE) { d->recs_h = h_chunk(d->recs_h, r, c, 1, 0); d->n_recs++; }
            else if (l->kind == ORC_L_DONE_EXACT) { d->recs_h = h_chunk(d->recs_h, r, 0, 1, 1); d->n_recs++; }
        } else {
            d->frames_h = h_str(d->frames_h, r->out + l->out_off, l->out_len); d->n_frames++; d->frame_bytes += l->out_len;
            if (c) { d->recs_h = h_chunk(d->recs_h, r, c, 0, 0); d->n_recs++; }
        }
    }
    *terminated = (uint8_t)((mode & 1u) ? r->terminated : 0);
    *tail_len = (uint32_t)r->tail_len;
}

typedef struct {
    const uint8_t *arena; const uint64_t *off; const uint32_t *len; const uint8_t *mode; size_t n;
    atomic_size_t *next; orc_digest *d; uint8_t *terminated; uint32_t *tail_len;
} cjob;

static void *cworker(void *p) {
    cjob *j = (cjob *)p;
    orc_result *r = orc_result_new();
    for (;;) {
        size_t s0 = atomic_fetch_add(j->next, 16);
        if (s0 >= j->n) break;
        size_t s1 = s0 + 16 < j->n ? s0 + 16 : j->n;
        for (size_t s = s0; s < s1; s++) digest_one(r, j->arena + j->off[s], j->len[s], j->mode[s], &j->d[s], &j->terminated[s], &j->tail_len[s]);
    }
    orc_result_free(r);
    return 0;
}

/* Oracle digests of n whole upstream bodies (mode bit0 R, bit1 parse). d must be orc_digest_init'ed. */
void orc_digest_streams(const uint8_t *arena, const uint64_t *off, const uint32_t *len, const uint8_t *mode, size_t n,
                        int n_threads, orc_digest *d, uint8_t *terminated, uint32_t *tail_len) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    atomic_size_t next = 0;
    cjob j = { arena, off, len, mode, n, &next, d, terminated, tail_len };
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], 0, cworker, &j);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], 0);
    free(th);
}

/* ---------------------------------------------------------------- GPU side */
static inline const uint8_t *at(const sse_result *res, uint32_t off) {      /* sse_at(), restated so that liborc does not link the product */
    return off >= res->in_base ? res->in + (off - res->in_base) : res->out + off;
}
static inline const uint8_t *span(const sse_result *res, uint32_t off, int in_text) { return in_text ? res->text + off : at(res, off); }

static uint64_t h_rec(uint64_t h, const sse_result *res, const sse_rec *rc, uint32_t *inexact) {
    const uint32_t fl = rc->flags;
    if (fl & (SSE_F_TOO_LONG | SSE_F_DEPTH_LIMIT)) (*inexact)++;
    h = h_u8(h, (fl & SSE_F_JSON_OK) != 0); h = h_u8(h, (fl & SSE_F_DONE_LINE) != 0); h = h_u8(h, (fl & SSE_F_DONE_EXACT) != 0);
    if (!(fl & SSE_F_JSON_OK)) return h;
    h = h_u32(h, rc->n_choices);
    h = h_u32(h, (fl & SSE_F_FINISH_MASK) >> SSE_F_FINISH_SHIFT);
    h = h_u8(h, (fl & SSE_F_HAS_USAGE) != 0);
    if (fl & SSE_F_HAS_USAGE) {
        const sse_usage *u = &res->usages[rc->usage];
        h = h_i64(h, u->prompt_tokens); h = h_i64(h, u->completion_tokens); h = h_i64(h, u->total_tokens);
    }
    h = h_str(h, span(res, rc->content_off, (fl & SSE_F_CONTENT_TEXT) != 0), rc->content_len);
    h = h_u8(h, (fl & SSE_F_TC_NONNIL) != 0); h = h_u8(h, (fl & SSE_F_TC_VALID) != 0);
    h = h_u32(h, rc->tc_count);
    uint32_t t = rc->tc_first;
    for (uint32_t k = 0; k < rc->tc_count; k++) {
        if (t == SSE_NONE || t >= res->n_tcs) { (*inexact) += 1u << 16; break; }      /* chain shorter than tc_count: never equal */
        const sse_tc *c = &res->tcs[t];
        h = h_i64(h, c->index);
        h = h_u8(h, (c->flags & SSE_TC_HAS_ID) != 0);
        if (c->flags & SSE_TC_HAS_ID) h = h_str(h, span(res, c->id_off, (c->flags & SSE_TC_ID_TEXT) != 0), c->id_len);
        h = h_u8(h, (c->flags & SSE_TC_HAS_TYPE) != 0);
        if (c->flags & SSE_TC_HAS_TYPE) h = h_str(h, span(res, c->type_off, (c->flags & SSE_TC_TYPE_TEXT) != 0), c->type_len);
        h = h_u8(h, (c->flags & SSE_TC_HAS_FUNC) != 0);
        h = h_str(h, span(res, c->name_off, (c->flags & SSE_TC_NAME_TEXT) != 0), c->name_len);
        h = h_str(h, span(res, c->args_off, (c->flags & SSE_TC_ARGS_TEXT) != 0), c->args_len);
        t = c->next;
    }
    return h;
}

/* Folds one batch result into the digests: segment i belongs to connection conn[i]. carry_len / seg_flags (per connection)
 * receive the segment's values (seg_flags OR-ed). Returns the number of segments folded. */
uint32_t orc_digest_result(const sse_result *res, const uint32_t *conn, orc_digest *d, uint32_t *carry_len, uint32_t *seg_flags) {
    for (uint32_t i = 0; i < res->n_segs; i++) {
        const sse_seg_result *sr = &res->segs[i];
        orc_digest *D = &d[conn[i]];
        const sse_run *run = &sr->run;
        for (;;) {
            for (uint32_t f = 0; f < run->frame_count; f++) {
                const sse_frame *fr = &res->frames[run->frame_first + f];
                D->frames_h = h_str(D->frames_h, at(res, fr->off), fr->len); D->n_frames++; D->frame_bytes += fr->len;
            }
            for (uint32_t k = 0; k < run->rec_count; k++) { D->recs_h = h_rec(D->recs_h, res, &res->recs[run->rec_first + k], &D->inexact); D->n_recs++; }
            if (run->next == SSE_NONE) break;
            run = &res->runs[run->next];
        }
        carry_len[conn[i]] = sr->carry_len;
        seg_flags[conn[i]] |= sr->flags;
    }
    return res->n_segs;
}
