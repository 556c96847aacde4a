// sse_device.cuh -- device-side data layout shared by the kernels and the host library.
//
// HBM layout (one context per GPU):
//   in      u8[in_arena_bytes]         micro-batch input: per-connection segments, 16-byte aligned
//   segs    sse_seg[max_segs]          segment descriptors (conn, in_off, in_len, mode)
//   conns   ConnState[max_conns]       persistent per-connection state (carry length, flags)
//   carry   u8[max_conns*carry_slot]   persistent held-back tail of each connection (provider.go:322
//                                      ReadBytes would keep these bytes inside bufio.Reader)
//   out, frames, recs, tcs, usages, text, runs, seg_results   per-batch results (see include/sse_gpu.h)
//   ctr     Counters                   bump allocators + work ticket, zeroed per launch
#pragma once
#include <cstddef>
#include <stdint.h>
#include "../../include/sse_gpu.h"

struct ConnState {
    uint32_t carry_len;
    uint32_t flags;
};
#define CONN_FINISHED 1u   // mode R: terminating chunk seen (agent.go:235-242); later bytes are never read
#define CONN_DEAD     2u   // line exceeded carry_slot_bytes
#define CONN_LONG     4u   // a line longer than the smem window is being assembled in the carry slot

#ifndef SSE_LEN_SHIFT
#define SSE_LEN_SHIFT 5
#endif
#define SSE_LEN_BUCKETS (4096 >> SSE_LEN_SHIFT)
#define SSE_N_BUCKETS (32 * SSE_LEN_BUCKETS)
struct alignas(8) Counters {
    uint32_t ticket;       // next segment to process
    uint32_t n_tcs;
    // two 8-byte pairs: the produce kernel allocates a round's output with ONE 64-bit atomicAdd per pair (low word | high word
    // << 32) instead of four 32-bit ones; every counter stays below 2^31 (capacities are 31-bit), so the low word never carries
    uint32_t out_bytes;
    uint32_t n_items;      // split pipeline: work items written by the produce kernel
    uint32_t n_frames;
    uint32_t n_recs;
    uint32_t n_usages;
    uint32_t text_bytes;
    uint32_t n_runs;
    int32_t  status;
    uint32_t item_ticket;  // split pipeline: next item batch for the decode kernel
    uint32_t n_tiles;      // fused pipeline: tiles written by the plan kernel
    uint32_t reserved1;
    uint32_t overflow;     // SSE_OVF_* bits: which arena was too small
    uint32_t pad[2];
    // device-only tail (the host reads the struct up to here)
    uint32_t class_count[SSE_N_BUCKETS];  // split pipeline: items per bucket (see item_bucket)
    uint32_t class_cursor[SSE_N_BUCKETS];
};

static_assert(offsetof(Counters, out_bytes) % 8 == 0 && offsetof(Counters, n_items) == offsetof(Counters, out_bytes) + 4 &&
              offsetof(Counters, n_frames) % 8 == 0 && offsetof(Counters, n_recs) == offsetof(Counters, n_frames) + 4,
              "the produce kernel adds to (out_bytes, n_items) and (n_frames, n_recs) with one 64-bit atomic each");

struct KParams {
    const uint8_t *in;
    const sse_seg *segs;
    uint32_t n_segs;
    uint32_t max_conns;
    ConnState *conns;
    uint8_t *carry;
    uint32_t carry_slot;
    // results
    uint8_t *out;            uint32_t cap_out;
    uint32_t in_base;        // in == out + in_base: arena offsets >= in_base are input bytes (zero-copy frames)
    sse_frame *frames;       uint32_t cap_frames;
    sse_rec *recs;           uint32_t cap_recs;
    sse_tc *tcs;             uint32_t cap_tcs;
    sse_usage *usages;       uint32_t cap_usages;
    uint8_t *text;           uint32_t cap_text;
    sse_run *runs;           uint32_t cap_runs;
    sse_seg_result *seg_results;
    Counters *ctr;
    // split pipeline (produce -> decode -> finalize)
    uint4 *items;            uint32_t cap_items;   // src, len | rmode << 31, rec, seg
    uint32_t *seg_term;                            // per segment: smallest terminating record index, SSE_NONE if none
    uint32_t flags;                                // sse_config.flags
    uint4 *items_sorted;                           // items ordered by (length bucket, shape class): the 32 lanes of a decode batch walk alike lines
    // fused pipeline (plan -> fused tile kernel)
    uint2 *tiles;            uint32_t cap_tiles;   // {first segment, segment count} per tile
    uint32_t *tcache;                              // skeleton templates kept between launches: [0] words used, [1..256] buckets, store
};

#ifdef __CUDACC__
__device__ __forceinline__ void sse_overflow(Counters *c, uint32_t which) {
    atomicOr(&c->overflow, which);
    atomicExch(&c->status, (int)SSE_ERR_OVERFLOW);
}
#endif

// launch wrappers (sse_kernel.cu)
int sse_launch_produce_kernel(const KParams &p, void *stream, int sm_count);           // split pipeline, stage 1
int sse_launch_decode_finalize(const KParams &p, void *stream, int sm_count, int device);  // split pipeline, stages 2+3
int sse_v2_prepare(int device);                                                        // builds + uploads the automaton tables
int sse_fused_prepare(int device);                                                     // fused pipeline: tables + kernel attributes
int sse_launch_fused(const KParams &p, void *stream, int sm_count, int device);        // plan kernel + fused tile kernel
uint32_t sse_fused_max_line(void);
uint32_t sse_fused_tcache_words(void);                                                 // size of KParams.tcache                                                     // longest carry_slot_bytes the fused kernel supports
