// sse_host.cu -- C ABI of libssegpu.so (include/sse_gpu.h): context, pinned staging, batch pipeline.
//
// One context per GPU. A batch slot owns pinned host staging + device buffers + a CUDA stream:
//   sse_submit:  H2D(segs, input)  ->  [wait: previous batch's kernel]  ->  kernel  ->  D2H(counters)
//   sse_collect: sync, then exact-size D2H of the result arrays, sync.
// Kernels of consecutive batches are ordered by an event because they share per-connection state
// (carry, finished flag) -- the per-connection FIFO of provider.go:307-340; H2D of batch k+1 and D2H of
// batch k overlap kernel k on their own streams.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>
#include "sse_device.cuh"

namespace {

thread_local char g_cuda_err[256] = "";

bool cu_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return false;
}
#define CU(call) do { if (!cu_ok((call), #call)) return SSE_ERR_CUDA; } while (0)

enum SlotState { SLOT_FREE = 0, SLOT_ACQUIRED, SLOT_SUBMITTED, SLOT_COLLECTED };

struct Slot {
    int state = SLOT_FREE;
    cudaStream_t stream = nullptr;
    cudaEvent_t kernel_done = nullptr;
    // host (pinned)
    uint8_t *h_in = nullptr; sse_seg *h_segs = nullptr;
    uint8_t *h_out = nullptr; sse_frame *h_frames = nullptr; sse_rec *h_recs = nullptr; sse_tc *h_tcs = nullptr;
    sse_usage *h_usages = nullptr; uint8_t *h_text = nullptr; sse_run *h_runs = nullptr; sse_seg_result *h_segres = nullptr;
    Counters *h_ctr = nullptr;
    // device
    uint8_t *d_in = nullptr; sse_seg *d_segs = nullptr;
    uint8_t *d_out = nullptr; sse_frame *d_frames = nullptr; sse_rec *d_recs = nullptr; sse_tc *d_tcs = nullptr;
    sse_usage *d_usages = nullptr; uint8_t *d_text = nullptr; sse_run *d_runs = nullptr; sse_seg_result *d_segres = nullptr;
    Counters *d_ctr = nullptr;
    uint4 *d_items = nullptr; uint32_t *d_segterm = nullptr;   // split pipeline scratch (device only)
    uint4 *d_items2 = nullptr;                                 // items ordered for the decode kernel
    uint2 *d_tiles = nullptr;                                  // fused pipeline: tile list of the plan kernel
    uint32_t n_segs = 0, in_bytes = 0;
};

} // namespace

// Device layout of a slot's byte arenas: [out arena | pad | in arena] in ONE allocation, so that a single base pointer and
// a 32-bit offset address both. Offsets >= in_base are input bytes (zero-copy frames and their spans).
static inline uint32_t in_base_of(const sse_config &cfg) { return (cfg.out_arena_bytes + 255u) & ~255u; }

struct sse_ctx {
    int device = 0;
    int sm_count = 0;
    sse_config cfg{};
    ConnState *d_conns = nullptr;
    uint8_t *d_carry = nullptr;
    uint32_t *d_tcache = nullptr;        // fused kernel: skeleton templates kept between launches
    cudaStream_t ctl_stream = nullptr;
    cudaEvent_t last_kernel = nullptr;   // completion of the most recently launched kernel
    bool have_last = false;
    std::vector<Slot> slots;
    uint64_t launches = 0;
};

namespace {

template <class T> bool dalloc(T *&p, size_t n) { return cu_ok(cudaMalloc((void **)&p, n ? n * sizeof(T) : sizeof(T)), "cudaMalloc"); }
template <class T> bool halloc(T *&p, size_t n) { return cu_ok(cudaHostAlloc((void **)&p, n ? n * sizeof(T) : sizeof(T), cudaHostAllocDefault), "cudaHostAlloc"); }

void free_slot(Slot &s) {
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.kernel_done) cudaEventDestroy(s.kernel_done);
    cudaFreeHost(s.h_in); cudaFreeHost(s.h_segs); cudaFreeHost(s.h_out); cudaFreeHost(s.h_frames); cudaFreeHost(s.h_recs);
    cudaFreeHost(s.h_tcs); cudaFreeHost(s.h_usages); cudaFreeHost(s.h_text); cudaFreeHost(s.h_runs); cudaFreeHost(s.h_segres);
    cudaFreeHost(s.h_ctr);
    cudaFree(s.d_segs); cudaFree(s.d_out);   // d_in lives inside d_out's allocation
    cudaFree(s.d_frames); cudaFree(s.d_recs); cudaFree(s.d_tcs);
    cudaFree(s.d_usages); cudaFree(s.d_text); cudaFree(s.d_runs); cudaFree(s.d_segres); cudaFree(s.d_ctr); cudaFree(s.d_items); cudaFree(s.d_segterm); cudaFree(s.d_items2); cudaFree(s.d_tiles);
}

KParams make_params(sse_ctx *c, Slot &s, uint32_t n_segs) {
    KParams p{};
    p.in = s.d_in; p.segs = s.d_segs; p.n_segs = n_segs; p.max_conns = c->cfg.max_conns;
    p.conns = c->d_conns; p.carry = c->d_carry; p.carry_slot = c->cfg.carry_slot_bytes;
    p.out = s.d_out; p.cap_out = c->cfg.out_arena_bytes; p.in_base = in_base_of(c->cfg);
    p.frames = s.d_frames; p.cap_frames = c->cfg.max_frames;
    p.recs = s.d_recs; p.cap_recs = c->cfg.max_recs;
    p.tcs = s.d_tcs; p.cap_tcs = c->cfg.max_tcs;
    p.usages = s.d_usages; p.cap_usages = c->cfg.max_usages;
    p.text = s.d_text; p.cap_text = c->cfg.text_arena_bytes;
    p.runs = s.d_runs; p.cap_runs = c->cfg.max_runs;
    p.seg_results = s.d_segres; p.ctr = s.d_ctr;
    p.items = s.d_items; p.cap_items = c->cfg.max_recs; p.seg_term = s.d_segterm;
    p.items_sorted = s.d_items2; p.flags = c->cfg.flags;
    p.tiles = s.d_tiles; p.cap_tiles = c->cfg.max_segs + 64;
    p.tcache = (c->cfg.flags & SSE_FLAG_TEMPLATES) ? c->d_tcache : nullptr;
    return p;
}

int validate_segs(sse_ctx *c, Slot &s, uint32_t n_segs, uint32_t in_bytes) {
    if (n_segs > c->cfg.max_segs || in_bytes > c->cfg.in_arena_bytes) return SSE_ERR_ARG;
    for (uint32_t i = 0; i < n_segs; i++) {
        const sse_seg &g = s.h_segs[i];
        if (g.conn >= c->cfg.max_conns || (g.in_off & 15u) || (uint64_t)g.in_off + g.in_len > in_bytes) return SSE_ERR_ARG;
    }
    return SSE_OK;
}

int do_upload(sse_ctx *c, Slot &s, uint32_t n_segs, uint32_t in_bytes, cudaStream_t st) {
    if (n_segs) CU(cudaMemcpyAsync(s.d_segs, s.h_segs, (size_t)n_segs * sizeof(sse_seg), cudaMemcpyHostToDevice, st));
    if (in_bytes) CU(cudaMemcpyAsync(s.d_in, s.h_in, ((size_t)in_bytes + 15) & ~(size_t)15, cudaMemcpyHostToDevice, st));
    s.n_segs = n_segs; s.in_bytes = in_bytes;
    (void)c;
    return SSE_OK;
}

int do_launch(sse_ctx *c, Slot &s, uint32_t n_segs, cudaStream_t st) {
    CU(cudaMemsetAsync(s.d_ctr, 0, sizeof(Counters), st));
    if (n_segs) {
        KParams p = make_params(c, s, n_segs);
        int e;
        if (c->cfg.flags & SSE_FLAG_KERNEL_FUSED) {   // plan + single-pass tile kernel
            e = sse_launch_fused(p, (void *)st, c->sm_count, c->device);
            c->launches += 2;
        }
        else {   // default: produce -> sort -> decode (templates / automaton) -> finalize
            e = sse_launch_produce_kernel(p, (void *)st, c->sm_count);
            if (e == 0) e = sse_launch_decode_finalize(p, (void *)st, c->sm_count, c->device);
            c->launches += 6;   // produce, bucket hist/scan/scatter, decode, finalize
        }
        if (e != 0) { cu_ok((cudaError_t)e, "stream kernel launch"); return SSE_ERR_CUDA; }
    }
    s.n_segs = n_segs;
    return SSE_OK;
}

int do_download(sse_ctx *c, Slot &s, sse_result *res, cudaStream_t st) {
    CU(cudaMemcpyAsync(s.h_ctr, s.d_ctr, offsetof(Counters, class_count), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const Counters k = *s.h_ctr;
    memset(res, 0, sizeof *res);
    res->status = k.status;
    res->n_segs = s.n_segs;
    if (k.status != SSE_OK) {   // overflow: the batch result is unusable, report loudly
        res->overflow = k.overflow;
        return k.status;
    }
    res->n_frames = k.n_frames; res->n_recs = k.n_recs; res->n_tcs = k.n_tcs; res->n_usages = k.n_usages;
    res->n_runs = k.n_runs; res->out_bytes = k.out_bytes; res->text_bytes = k.text_bytes;
    res->n_decoded = k.n_items; res->n_derived = 0;
    if (k.out_bytes) CU(cudaMemcpyAsync(s.h_out, s.d_out, k.out_bytes, cudaMemcpyDeviceToHost, st));
    if (k.n_frames) CU(cudaMemcpyAsync(s.h_frames, s.d_frames, (size_t)k.n_frames * sizeof(sse_frame), cudaMemcpyDeviceToHost, st));
    if (k.n_recs) CU(cudaMemcpyAsync(s.h_recs, s.d_recs, (size_t)k.n_recs * sizeof(sse_rec), cudaMemcpyDeviceToHost, st));
    if (k.n_tcs) CU(cudaMemcpyAsync(s.h_tcs, s.d_tcs, (size_t)k.n_tcs * sizeof(sse_tc), cudaMemcpyDeviceToHost, st));
    if (k.n_usages) CU(cudaMemcpyAsync(s.h_usages, s.d_usages, (size_t)k.n_usages * sizeof(sse_usage), cudaMemcpyDeviceToHost, st));
    if (k.text_bytes) CU(cudaMemcpyAsync(s.h_text, s.d_text, k.text_bytes, cudaMemcpyDeviceToHost, st));
    if (k.n_runs) CU(cudaMemcpyAsync(s.h_runs, s.d_runs, (size_t)k.n_runs * sizeof(sse_run), cudaMemcpyDeviceToHost, st));
    if (s.n_segs) CU(cudaMemcpyAsync(s.h_segres, s.d_segres, (size_t)s.n_segs * sizeof(sse_seg_result), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    res->out = s.h_out; res->frames = s.h_frames; res->recs = s.h_recs; res->tcs = s.h_tcs; res->usages = s.h_usages;
    res->text = s.h_text; res->runs = s.h_runs; res->segs = s.h_segres;
    res->in = s.h_in; res->in_base = in_base_of(c->cfg);
    return SSE_OK;
}

} // namespace

extern "C" {

int sse_abi_version(void) { return SSE_ABI_VERSION; }

const uint8_t *sse_at(const sse_result *r, uint32_t off) { return off >= r->in_base ? r->in + (off - r->in_base) : r->out + off; }

const char *sse_last_cuda_error(void) { return g_cuda_err; }

const char *sse_strerror(int status) {
    switch (status) {
    case SSE_OK: return "ok";
    case SSE_ERR_NO_DEVICE: return "no CUDA device available (libssegpu has no CPU fallback)";
    case SSE_ERR_CUDA: return "CUDA error (see sse_last_cuda_error)";
    case SSE_ERR_ARG: return "invalid argument";
    case SSE_ERR_BUSY: return "batch slot busy or in the wrong state";
    case SSE_ERR_OVERFLOW: return "result arena overflow: batch discarded (increase sse_config capacities)";
    case SSE_ERR_NOMEM: return "out of memory";
    case SSE_ERR_UNDECODED: return "a line was not decoded (SSE_F_TOO_LONG / SSE_F_DEPTH_LIMIT): fold results are incomplete";
    default: return "unknown status";
    }
}

void sse_default_config(sse_config *cfg, uint32_t max_conns, uint32_t bytes_per_batch) {
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = sizeof *cfg;
    cfg->max_conns = max_conns;
    cfg->max_segs = max_conns;
    uint64_t in = ((uint64_t)bytes_per_batch + 16ull * max_conns + 4095) & ~4095ull;
    cfg->in_arena_bytes = (uint32_t)(in > 0xF0000000ull ? 0xF0000000ull : in);
    uint64_t out = in + in / 8 + 65536;
    cfg->out_arena_bytes = (uint32_t)(out > 0xF0000000ull ? 0xF0000000ull : out);
    cfg->max_frames = (uint32_t)(in / 32 + 4ull * max_conns + 1024);
    cfg->max_recs = (uint32_t)(in / 64 + 2ull * max_conns + 1024);
    cfg->max_tcs = (uint32_t)(in / 256 + 1024);
    cfg->max_usages = cfg->max_recs;
    cfg->text_arena_bytes = (uint32_t)(in / 4 + 65536);
    cfg->max_runs = (uint32_t)(in / 2048 + max_conns + 1024);   // one extra run per window / 64 lines of a long segment
    cfg->carry_slot_bytes = 16384;
    cfg->n_slots = 2;
}

void sse_worst_case_config(sse_config *cfg, uint32_t max_conns, uint32_t bytes_per_batch) {
    sse_default_config(cfg, max_conns, bytes_per_batch);
    const uint64_t in = cfg->in_arena_bytes, cap = 0xF0000000ull;
    auto clamp = [&](uint64_t v) { return (uint32_t)(v > cap ? cap : v); };
    cfg->carry_slot_bytes = 65536;                                   // longest supported line
    cfg->max_frames = clamp(in + 64);                                // a frame needs its own '\n' in this batch
    cfg->max_recs = clamp(in / 4 + 2ull * max_conns + 64);           // "data: x\n" is 8 bytes; one carried line per segment
    cfg->max_usages = cfg->max_recs;
    cfg->max_tcs = clamp(in / 3 + 64);                               // "{}," per element
    cfg->text_arena_bytes = clamp(4 * in + 65536);                   // U+FFFD for every invalid byte: 1 -> 3, 4-byte aligned
    cfg->out_arena_bytes = clamp(in + in / 4 + (uint64_t)max_conns * (cfg->carry_slot_bytes + 32ull) + (1u << 20));
    cfg->max_runs = clamp(in / 256 + 2ull * max_conns + 1024);
}

int sse_init(int device, const sse_config *cfg, sse_ctx **out) {
    if (!cfg || !out || cfg->struct_size != sizeof(sse_config)) return SSE_ERR_ARG;
    if (cfg->n_slots < 1 || cfg->n_slots > 8 || cfg->max_conns == 0 || cfg->max_segs == 0 ||
        (cfg->in_arena_bytes & 15u) || cfg->carry_slot_bytes < 8192 + 16 || (cfg->carry_slot_bytes & 15u)) return SSE_ERR_ARG;
    if ((uint64_t)in_base_of(*cfg) + cfg->in_arena_bytes + 16 >= (1ull << 31)) return SSE_ERR_ARG;   // arena offsets are 31-bit
    if (cfg->flags & ~(SSE_FLAG_KERNEL_FUSED | SSE_FLAG_COPY_OUT | SSE_FLAG_TEMPLATES)) return SSE_ERR_ARG;          // unknown engine flag
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
        cu_ok(cudaGetLastError(), "cudaGetDeviceCount");
        return SSE_ERR_NO_DEVICE;
    }
    CU(cudaSetDevice(device));
    sse_ctx *c = new (std::nothrow) sse_ctx();
    if (!c) return SSE_ERR_NOMEM;
    c->device = device; c->cfg = *cfg;
    cudaDeviceProp prop;
    if (!cu_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) { delete c; return SSE_ERR_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    if (cfg->flags & SSE_FLAG_KERNEL_FUSED) {
        if (cfg->carry_slot_bytes > sse_fused_max_line()) { delete c; return SSE_ERR_ARG; }
        int e2 = sse_fused_prepare(device);
        if (e2 != 0) { cu_ok((cudaError_t)e2, "sse_fused_prepare"); delete c; return SSE_ERR_CUDA; }
    } else {
        int e2 = sse_v2_prepare(device);
        if (e2 != 0) { cu_ok((cudaError_t)e2, "sse_v2_prepare"); delete c; return SSE_ERR_CUDA; }
    }
    bool ok = true;
    ok = ok && cu_ok(cudaStreamCreateWithFlags(&c->ctl_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    ok = ok && cu_ok(cudaEventCreateWithFlags(&c->last_kernel, cudaEventDisableTiming), "cudaEventCreate");
    ok = ok && dalloc(c->d_conns, cfg->max_conns);
    ok = ok && dalloc(c->d_carry, (size_t)cfg->max_conns * cfg->carry_slot_bytes);
    ok = ok && cu_ok(cudaMemset(c->d_conns, 0, (size_t)cfg->max_conns * sizeof(ConnState)), "cudaMemset");
    ok = ok && dalloc(c->d_tcache, sse_fused_tcache_words());
    ok = ok && cu_ok(cudaMemset(c->d_tcache, 0, (size_t)sse_fused_tcache_words() * sizeof(uint32_t)), "cudaMemset");
    c->slots.resize(cfg->n_slots);
    for (auto &s : c->slots) {
        if (!ok) break;
        ok = ok && cu_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        ok = ok && cu_ok(cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming), "cudaEventCreate");
        ok = ok && halloc(s.h_in, (size_t)cfg->in_arena_bytes + 16) && halloc(s.h_segs, cfg->max_segs);
        ok = ok && halloc(s.h_out, cfg->out_arena_bytes) && halloc(s.h_frames, cfg->max_frames) && halloc(s.h_recs, cfg->max_recs);
        ok = ok && halloc(s.h_tcs, cfg->max_tcs) && halloc(s.h_usages, cfg->max_usages) && halloc(s.h_text, cfg->text_arena_bytes);
        ok = ok && halloc(s.h_runs, cfg->max_runs) && halloc(s.h_segres, cfg->max_segs) && halloc(s.h_ctr, 1);
        ok = ok && dalloc(s.d_segs, cfg->max_segs);
        ok = ok && dalloc(s.d_out, (size_t)in_base_of(*cfg) + cfg->in_arena_bytes + 16);
        if (ok) s.d_in = s.d_out + in_base_of(*cfg);
        ok = ok && dalloc(s.d_frames, cfg->max_frames) && dalloc(s.d_recs, cfg->max_recs);
        ok = ok && dalloc(s.d_tcs, cfg->max_tcs) && dalloc(s.d_usages, cfg->max_usages) && dalloc(s.d_text, cfg->text_arena_bytes);
        ok = ok && dalloc(s.d_runs, cfg->max_runs) && dalloc(s.d_segres, cfg->max_segs) && dalloc(s.d_ctr, 1);
        ok = ok && dalloc(s.d_items, cfg->max_recs) && dalloc(s.d_segterm, cfg->max_segs);
        ok = ok && dalloc(s.d_items2, cfg->max_recs);
        ok = ok && dalloc(s.d_tiles, (size_t)cfg->max_segs + 64);
    }
    if (!ok) { sse_destroy(c); return SSE_ERR_CUDA; }
    *out = c;
    return SSE_OK;
}

void sse_destroy(sse_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &s : c->slots) free_slot(s);
    cudaFree(c->d_conns); cudaFree(c->d_carry); cudaFree(c->d_tcache);
    if (c->ctl_stream) cudaStreamDestroy(c->ctl_stream);
    if (c->last_kernel) cudaEventDestroy(c->last_kernel);
    delete c;
}

int sse_acquire(sse_ctx *c, int *slot, sse_batch *batch) {
    if (!c || !slot || !batch) return SSE_ERR_ARG;
    for (size_t i = 0; i < c->slots.size(); i++) {
        if (c->slots[i].state == SLOT_FREE) {
            Slot &s = c->slots[i];
            s.state = SLOT_ACQUIRED;
            *slot = (int)i;
            batch->in_arena = s.h_in; batch->segs = s.h_segs;
            batch->in_arena_bytes = c->cfg.in_arena_bytes; batch->max_segs = c->cfg.max_segs;
            return SSE_OK;
        }
    }
    return SSE_ERR_BUSY;
}

int sse_submit(sse_ctx *c, int slot, uint32_t n_segs, uint32_t in_bytes) {
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return SSE_ERR_ARG;
    Slot &s = c->slots[slot];
    if (s.state != SLOT_ACQUIRED) return SSE_ERR_BUSY;
    int rc = validate_segs(c, s, n_segs, in_bytes);
    if (rc != SSE_OK) return rc;
    CU(cudaSetDevice(c->device));
    rc = do_upload(c, s, n_segs, in_bytes, s.stream);
    if (rc != SSE_OK) return rc;
    if (c->have_last) CU(cudaStreamWaitEvent(s.stream, c->last_kernel, 0));   // per-connection FIFO across batches
    rc = do_launch(c, s, n_segs, s.stream);
    if (rc != SSE_OK) return rc;
    CU(cudaEventRecord(c->last_kernel, s.stream));
    c->have_last = true;
    s.state = SLOT_SUBMITTED;
    return SSE_OK;
}

int sse_collect(sse_ctx *c, int slot, sse_result *res) {
    if (!c || !res || slot < 0 || slot >= (int)c->slots.size()) return SSE_ERR_ARG;
    Slot &s = c->slots[slot];
    if (s.state != SLOT_SUBMITTED) return SSE_ERR_BUSY;
    CU(cudaSetDevice(c->device));
    int rc = do_download(c, s, res, s.stream);
    s.state = SLOT_COLLECTED;
    return rc;
}

int sse_release(sse_ctx *c, int slot) {
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return SSE_ERR_ARG;
    Slot &s = c->slots[slot];
    if (s.state == SLOT_SUBMITTED) { CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(s.stream)); }
    s.state = SLOT_FREE;
    return SSE_OK;
}

int sse_reset_conn(sse_ctx *c, uint32_t conn) {
    if (!c || conn >= c->cfg.max_conns) return SSE_ERR_ARG;
    CU(cudaSetDevice(c->device));
    if (c->have_last) CU(cudaStreamWaitEvent(c->ctl_stream, c->last_kernel, 0));
    CU(cudaMemsetAsync(c->d_conns + conn, 0, sizeof(ConnState), c->ctl_stream));
    CU(cudaEventRecord(c->last_kernel, c->ctl_stream));   // later kernels are ordered after the reset
    c->have_last = true;
    return SSE_OK;
}

int sse_reset_all(sse_ctx *c, void *cuda_stream) {
    if (!c) return SSE_ERR_ARG;
    CU(cudaSetDevice(c->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->ctl_stream;
    if (!cuda_stream && c->have_last) CU(cudaStreamWaitEvent(st, c->last_kernel, 0));
    CU(cudaMemsetAsync(c->d_conns, 0, (size_t)c->cfg.max_conns * sizeof(ConnState), st));
    if (!cuda_stream) { CU(cudaEventRecord(c->last_kernel, st)); c->have_last = true; }
    return SSE_OK;
}

int sse_upload(sse_ctx *c, int slot, uint32_t n_segs, uint32_t in_bytes, void *cuda_stream) {
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return SSE_ERR_ARG;
    Slot &s = c->slots[slot];
    if (s.state != SLOT_ACQUIRED) return SSE_ERR_BUSY;
    int rc = validate_segs(c, s, n_segs, in_bytes);
    if (rc != SSE_OK) return rc;
    CU(cudaSetDevice(c->device));
    return do_upload(c, s, n_segs, in_bytes, cuda_stream ? (cudaStream_t)cuda_stream : s.stream);
}

int sse_launch(sse_ctx *c, int slot, uint32_t n_segs, void *cuda_stream) {
    if (!c || slot < 0 || slot >= (int)c->slots.size()) return SSE_ERR_ARG;
    Slot &s = c->slots[slot];
    if (s.state != SLOT_ACQUIRED || n_segs > c->cfg.max_segs) return SSE_ERR_BUSY;
    CU(cudaSetDevice(c->device));
    return do_launch(c, s, n_segs, cuda_stream ? (cudaStream_t)cuda_stream : s.stream);
}

int sse_download(sse_ctx *c, int slot, sse_result *res, void *cuda_stream) {
    if (!c || !res || slot < 0 || slot >= (int)c->slots.size()) return SSE_ERR_ARG;
    Slot &s = c->slots[slot];
    if (s.state != SLOT_ACQUIRED) return SSE_ERR_BUSY;
    CU(cudaSetDevice(c->device));
    return do_download(c, s, res, cuda_stream ? (cudaStream_t)cuda_stream : s.stream);
}

/* debugging aid (tools/dump_templates.py): the fused kernel's template cache, as sse_fused.cu lays it out */
int sse_debug_tcache(sse_ctx *c, uint32_t *out, uint32_t words) {
    if (!c || !out || !c->d_tcache) return SSE_ERR_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    const uint32_t n = words < sse_fused_tcache_words() ? words : sse_fused_tcache_words();
    CU(cudaMemcpy(out, c->d_tcache, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return (int)n;
}

int sse_launch_count(sse_ctx *c, uint64_t *kernel_launches) {
    if (!c || !kernel_launches) return SSE_ERR_ARG;
    *kernel_launches = c->launches;
    return SSE_OK;
}

} // extern "C"
