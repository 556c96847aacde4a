/*
 * sse_gpu.h -- C ABI of libssegpu.so: the B200 (sm_100a) implementation of inference-gateway's
 * streaming-response hot path.  This is the drop-in boundary: the Go side (integration/go, see
 * INTEGRATION.md) binds exactly these symbols through cgo and keeps core.IProvider /
 * registry.ProviderRegistry unchanged above them.
 *
 * What each entry point replaces in the reference (paths relative to the reference root):
 *   sse_submit/sse_collect   the per-line work of ProviderImpl.StreamChatCompletions' reader goroutine
 *                            (providers/core/provider.go:308-341: ReadBytes('\n'), tail held back),
 *                            the verbatim writers (api/routes.go:600-625, :178-231) [mode P], and the
 *                            MCP agent's per-line trim / [DONE] / "data: " / reframe / json.Unmarshal /
 *                            early-termination loop (mcp/agent.go:169-248) [mode R]
 *   sse_rec / sse_tc / sse_usage   the decoded fields of types.CreateChatCompletionStreamResponse that
 *                            the reference consumes (providers/types/common_types.go:271-297,:300-346,
 *                            :384-393,:451-478), i.e. what agent.go:199-242, agent.go:377-481 and
 *                            api/middlewares/telemetry.go:190-277 read after json.Unmarshal
 *   sse_reset_conn           a new upstream stream on the slot (new ProviderImpl per request,
 *                            providers/registry/registry.go:58-69; next agent iteration, agent.go:145-148)
 *   sse_agent_* / sse_telemetry_*  host-side folds of the records that reproduce the accumulators of
 *                            agent.go:156-260 (+ :377-481) and telemetry.go:190-277
 *
 * Conventions: every function returns 0 (SSE_OK) or a negative sse_status; nothing throws across the
 * ABI; every buffer handed to the caller is allocated by the library (cudaHostAlloc) and stays valid
 * until sse_release()/sse_destroy(); the library never retains caller pointers.  There is NO CPU
 * fallback: without a CUDA device sse_init fails with SSE_ERR_NO_DEVICE.
 */
#ifndef SSE_GPU_H
#define SSE_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSE_ABI_VERSION 2

typedef enum {
    SSE_OK = 0,
    SSE_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path has no CPU fallback */
    SSE_ERR_CUDA = -2,        /* a CUDA call failed; sse_last_cuda_error() has the text */
    SSE_ERR_ARG = -3,
    SSE_ERR_BUSY = -4,        /* no free batch slot / slot in wrong state */
    SSE_ERR_OVERFLOW = -5,    /* a result arena was too small for this batch (config too small) */
    SSE_ERR_NOMEM = -6,
    SSE_ERR_UNDECODED = -7    /* sse_agent_feed / sse_telemetry_feed: the segment holds a record flagged SSE_F_TOO_LONG or SSE_F_DEPTH_LIMIT, i.e. a
                                 line the reference would have decoded and this library did not: the fold took the rest, the caller must
                                 fail the stream (or decode that frame itself) instead of trusting the accumulators */
} sse_status;

/* segment modes (per connection, chosen by which reference consumer sits above the provider) */
#define SSE_MODE_P     0u  /* passthrough: every complete line verbatim (routes.go:613, :220) */
#define SSE_MODE_R     1u  /* MCP reframe "data: "+payload+"\n\n" with agent.go:178-242 rules; implies PARSE */
#define SSE_MODE_PARSE 2u  /* decode payloads of lines that start with "data: " into sse_rec (telemetry tap) */

typedef struct {
    uint32_t struct_size;      /* sizeof(sse_config), for ABI evolution */
    uint32_t max_conns;        /* connection slots [0, max_conns) with persistent carry state */
    uint32_t max_segs;         /* segments per batch (at most one per connection per batch) */
    uint32_t in_arena_bytes;   /* input bytes per batch */
    uint32_t out_arena_bytes;  /* emitted bytes per batch */
    uint32_t max_frames;       /* emitted frames per batch */
    uint32_t max_recs;         /* parsed lines per batch */
    uint32_t max_tcs;          /* tool-call elements per batch */
    uint32_t max_usages;       /* usage objects per batch */
    uint32_t text_arena_bytes; /* decoded (unescaped) strings per batch */
    uint32_t max_runs;         /* extra result runs per batch (segments with > 64 lines) */
    uint32_t carry_slot_bytes; /* per-connection held-back tail capacity == longest supported line */
    uint32_t n_slots;          /* batches in flight (pipeline depth), 1..8 */
    uint32_t flags;            /* SSE_FLAG_* */
} sse_config;

#define SSE_FLAG_KERNEL_FUSED 4u /* single-pass tile kernel (1-D TMA staging, stop bitmap, warp-per-line template replay): reads the payload from HBM
                                    once; slower than the default pipeline on B200 */
#define SSE_FLAG_TEMPLATES 16u /* skeleton-template replay: lines whose JSON skeleton was seen before skip the automaton (results identical;
                                  measured slower than the automaton alone on B200, DESIGN.md 4.4, so off by default) */
#define SSE_FLAG_COPY_OUT 8u   /* materialise every frame in the out arena. Default: a frame whose bytes already stand in the
                                  caller's input arena exactly as the reference would send them (every mode P line, and a mode R
                                  "data: ...\n" line followed by a blank line) is returned as a span of the input arena and is
                                  neither copied on the device nor sent back over PCIe (see sse_at) */

/* One segment = the bytes read from ONE connection since the previous batch. in_off is 16-byte aligned. */
typedef struct {
    uint32_t conn;
    uint32_t in_off;
    uint32_t in_len;
    uint8_t  mode;             /* SSE_MODE_* bits */
    uint8_t  provider;         /* remap-table selector; all 11 reference providers are identity */
    uint16_t reserved;
} sse_seg;

/* An emitted frame: one element of the reference's chan []byte. `off` is an arena offset: below sse_result.in_base it
 * indexes sse_result.out, at or above it indexes the batch's own input arena (off - in_base); use sse_at(). The same
 * holds for every span in sse_rec / sse_tc that is not flagged *_TEXT. */
typedef struct { uint32_t off, len; } sse_frame;

/* sse_rec.flags */
#define SSE_F_JSON_OK        0x0001u /* json.Unmarshal returned nil (syntax valid, no type mismatch) */
#define SSE_F_HAS_USAGE      0x0002u /* resp.Usage != nil */
#define SSE_F_TC_NONNIL      0x0004u /* choices[0].delta.tool_calls != nil */
#define SSE_F_TC_VALID       0x0008u /* agent.go:224-233 predicate true for some element */
#define SSE_F_CONTENT_TEXT   0x0010u /* content span is in the text arena (needed unescaping), else an arena offset (sse_at) */
#define SSE_F_DONE_LINE      0x0020u /* mode R: line contained "[DONE]" and was swallowed (agent.go:181-184) */
#define SSE_F_DONE_EXACT     0x0040u /* ... and its payload is exactly "[DONE]" (agent.go:394-396 break) */
#define SSE_F_TERMINATES     0x0080u /* mode R: finish_reason stop/tool_calls on an emitted chunk (agent.go:235-242) */
#define SSE_F_DEPTH_LIMIT    0x0100u /* nesting deeper than 128: reported as not JSON_OK (documented limit) */
#define SSE_F_TOO_LONG       0x0200u /* line longer than the parse window: frame exact, side-band not decoded */
#define SSE_F_FINISH_SHIFT   12      /* bits 12..14: SSE_FIN_* of choices[0].finish_reason */
#define SSE_F_FINISH_MASK    0x7000u
enum { SSE_FIN_NONE = 0, SSE_FIN_STOP = 1, SSE_FIN_TOOL_CALLS = 2, SSE_FIN_LENGTH = 3,
       SSE_FIN_CONTENT_FILTER = 4, SSE_FIN_FUNCTION_CALL = 5, SSE_FIN_OTHER = 7 };

#define SSE_NONE 0xFFFFFFFFu

/* One record per line whose payload was decoded. 32 bytes. */
typedef struct {
    uint32_t frame;        /* index into frames[] of the emitted frame, SSE_NONE if not emitted */
    uint32_t flags;        /* SSE_F_* */
    uint32_t content_off;  /* decoded choices[0].delta.content */
    uint32_t content_len;
    uint32_t tc_first;     /* first tool-call element (linked through sse_tc.next), SSE_NONE if none */
    uint16_t tc_count;     /* len(*choices[0].delta.tool_calls), saturating */
    uint16_t n_choices;    /* len(resp.Choices), saturating */
    uint32_t usage;        /* index into usages[], SSE_NONE if resp.Usage == nil */
    uint32_t payload_len;  /* bytes handed to json.Unmarshal */
} sse_rec;

/* sse_tc.flags */
#define SSE_TC_HAS_ID    0x01u
#define SSE_TC_HAS_TYPE  0x02u
#define SSE_TC_HAS_FUNC  0x04u
#define SSE_TC_ID_TEXT   0x10u   /* span is in the text arena (else an arena offset, sse_at) */
#define SSE_TC_TYPE_TEXT 0x20u
#define SSE_TC_NAME_TEXT 0x40u
#define SSE_TC_ARGS_TEXT 0x80u

/* One element of choices[0].delta.tool_calls (ChatCompletionMessageToolCallChunk). 48 bytes. */
typedef struct {
    int64_t  index;
    uint32_t flags;
    uint32_t next;         /* next element of the same chunk, SSE_NONE at the end */
    uint32_t id_off, id_len, type_off, type_len, name_off, name_len, args_off, args_len;
} sse_tc;

typedef struct { int64_t prompt_tokens, completion_tokens, total_tokens; } sse_usage;

/* sse_seg_result.flags */
#define SSE_SEG_TERMINATED   0x01u  /* a chunk with SSE_F_TERMINATES was emitted in this batch */
#define SSE_SEG_FINISHED     0x02u  /* connection had terminated earlier: bytes ignored (never read by the reference) */
#define SSE_SEG_LINE_TOO_LONG 0x04u /* a line exceeded carry_slot_bytes: stream failed, connection dead until reset */
#define SSE_SEG_DEAD         0x08u  /* connection is in the failed state */

/* A run: contiguous frames/recs produced by up to 64 consecutive lines of one segment. */
typedef struct {
    uint32_t frame_first, frame_count;
    uint32_t rec_first, rec_count;
    uint32_t next;         /* next run of the same segment in runs[], SSE_NONE at the end */
} sse_run;

/* Per-segment result, same index as the submitted sse_seg. 32 bytes. */
typedef struct {
    sse_run  run;          /* first run inline */
    uint32_t carry_len;    /* bytes held back (unterminated tail), carried to the next batch */
    uint32_t flags;        /* SSE_SEG_* */
    uint32_t reserved;
} sse_seg_result;

typedef struct {
    int32_t  status;       /* SSE_OK or SSE_ERR_OVERFLOW */
    uint32_t n_segs, n_frames, n_recs, n_tcs, n_usages, n_runs;
    uint32_t out_bytes, text_bytes;
    const uint8_t        *out;      /* emitted bytes that had to be materialised (reframed, carried over or assembled lines) */
    const sse_frame      *frames;
    const sse_rec        *recs;
    const sse_tc         *tcs;
    const sse_usage      *usages;
    const uint8_t        *text;     /* decoded strings */
    const sse_run        *runs;
    const sse_seg_result *segs;
    uint32_t n_decoded, n_derived;  /* statistics: lines decoded by the automaton; n_derived is reserved (0) */
    uint32_t overflow;              /* SSE_OVF_* bits when status == SSE_ERR_OVERFLOW: which sse_config capacity to raise */
    uint32_t in_base;               /* arena offsets >= in_base refer to in[off - in_base] */
    const uint8_t *in;              /* the batch's input arena (sse_batch.in_arena), lent until sse_release like out */
} sse_result;

/* Resolve an arena offset (frame.off, content_off, tool-call spans without the *_TEXT flag):
 * off >= r->in_base ? r->in + (off - r->in_base) : r->out + off. */
const uint8_t *sse_at(const sse_result *r, uint32_t off);

#define SSE_OVF_OUT    0x01u  /* out_arena_bytes / max_frames / max_recs (bump-allocated together per round) */
#define SSE_OVF_TCS    0x02u  /* max_tcs */
#define SSE_OVF_USAGES 0x04u  /* max_usages */
#define SSE_OVF_TEXT   0x08u  /* text_arena_bytes */
#define SSE_OVF_RUNS   0x10u  /* max_runs */

typedef struct {
    uint8_t *in_arena;     /* pinned; caller writes each segment's bytes at a 16-byte aligned in_off */
    sse_seg *segs;         /* pinned; caller writes descriptors */
    uint32_t in_arena_bytes, max_segs;
} sse_batch;

typedef struct sse_ctx sse_ctx;

/* lifecycle */
int  sse_init(int device, const sse_config *cfg, sse_ctx **out);
void sse_destroy(sse_ctx *ctx);
const char *sse_strerror(int status);
const char *sse_last_cuda_error(void);
int  sse_abi_version(void);
void sse_default_config(sse_config *cfg, uint32_t max_conns, uint32_t bytes_per_batch);
/* Capacities under which NO input can overflow a result arena (a batch of bytes_per_batch bytes on max_conns connections):
 * memory grows to ~30x the batch size, meant for small batcher arenas (sse_gateway.h uses it). With sse_default_config the
 * capacities fit realistic SSE traffic and a pathological batch fails as a whole with SSE_ERR_OVERFLOW. */
void sse_worst_case_config(sse_config *cfg, uint32_t max_conns, uint32_t bytes_per_batch);

/* pipeline: acquire -> fill -> submit -> collect -> release (one slot = one batch in flight) */
int sse_acquire(sse_ctx *ctx, int *slot, sse_batch *batch);
int sse_submit(sse_ctx *ctx, int slot, uint32_t n_segs, uint32_t in_bytes);
int sse_collect(sse_ctx *ctx, int slot, sse_result *res);
int sse_release(sse_ctx *ctx, int slot);

/* connection state */
int sse_reset_conn(sse_ctx *ctx, uint32_t conn);                 /* ordered after already submitted batches */
int sse_reset_all(sse_ctx *ctx, void *cuda_stream);              /* all connections; async on the stream */

/* device-resident replay (bench / roofline measurement): the three stages of sse_submit/sse_collect */
int sse_upload(sse_ctx *ctx, int slot, uint32_t n_segs, uint32_t in_bytes, void *cuda_stream);
int sse_launch(sse_ctx *ctx, int slot, uint32_t n_segs, void *cuda_stream);   /* kernels only */
int sse_download(sse_ctx *ctx, int slot, sse_result *res, void *cuda_stream); /* D2H + sync */
int sse_launch_count(sse_ctx *ctx, uint64_t *kernel_launches);                /* kernels launched so far */

/* ---- host-side folds of the records (the reference's consumers, fed from the side-band) ---- */
typedef struct { const uint8_t *p; size_t n; } sse_bytes;
typedef struct { sse_bytes id, type, name, arguments; } sse_tool_call;

/* mcp/agent.go:156-260 for one agent iteration: content accumulator, hasToolCalls, finish, and
 * parseStreamingToolCalls (:377-481) evaluated from records instead of re-parsing the text. */
typedef struct sse_agent_fold sse_agent_fold;
sse_agent_fold *sse_agent_new(void);
void sse_agent_free(sse_agent_fold *f);
void sse_agent_reset(sse_agent_fold *f);
int  sse_agent_feed(sse_agent_fold *f, const sse_result *res, uint32_t seg_index);
sse_bytes sse_agent_content(const sse_agent_fold *f);
int  sse_agent_has_tool_calls(const sse_agent_fold *f);
int  sse_agent_terminated(const sse_agent_fold *f, int *finish);
size_t sse_agent_tool_calls(sse_agent_fold *f, sse_tool_call *calls, size_t cap);

/* api/middlewares/telemetry.go:190-277 evaluated from records of a mode P|PARSE or mode R stream. */
typedef struct sse_telemetry_fold sse_telemetry_fold;
sse_telemetry_fold *sse_telemetry_new(void);
void sse_telemetry_free(sse_telemetry_fold *f);
void sse_telemetry_reset(sse_telemetry_fold *f);
int  sse_telemetry_feed(sse_telemetry_fold *f, const sse_result *res, uint32_t seg_index);
/* Bytes the HOST writes to the client between upstream frames: ssegw_agent_recv's final "data: [DONE]\n\n" (agent.go:140-143).
 * They are part of the body telemetry.go:190-198 splits into pieces, so they move its last-4-pieces window; they are not decoded
 * (feed only frames that hold no chunk: "[DONE]", comments, error text). n must end on a '\n'. */
int  sse_telemetry_feed_bytes(sse_telemetry_fold *f, const uint8_t *bytes, size_t n);
int  sse_telemetry_finish(sse_telemetry_fold *f, sse_usage *usage, sse_tool_call *calls, size_t cap, size_t *n_calls);

#ifdef __cplusplus
}
#endif
#endif
