// sse_gateway.cpp -- host-side mirror of the reference's interfaces for the streaming path, above the C ABI.
//
// The reference is compiled Go; its toolchain is absent here, so the host side is C++ with the reference's names and
// error behaviour (the Go shim in integration/go/ is the same logic over cgo):
//
//   ssegw_stream_chat_completions  ~ core.IProvider.StreamChatCompletions (providers/core/interfaces.go:22,
//                                    provider.go:277-344): returns a stream whose channel carries one []byte per line
//   ssegw_upstream_write/close     ~ bytes arriving on response.Body / EOF-or-error (provider.go:322-330: the
//                                    unterminated tail is dropped, then the channel is closed)
//   ssegw_pump                     ~ one tick of the per-GPU batcher goroutine: micro-batch every
//                                    stream's pending bytes, sse_submit, sse_collect, deliver frames to the channels
//   ssegw_recv                     ~ `line, ok := <-streamCh` (api/routes.go:602-606, mcp/agent.go:171)
//   ssegw_agent_*                  ~ mcp.Agent.RunWithStream for one iteration (mcp/agent.go:126-290): forwards frames to
//                                    the middleware channel, exposes accumulated content / tool calls / finish, appends the
//                                    single final "data: [DONE]\n\n" (agent.go:140-143)
//
// Channel capacity is 100 as in provider.go:307: a stream with 100 undelivered elements stops accepting upstream bytes
// (ssegw_upstream_write returns 0 bytes accepted), which is the back-pressure the reader goroutine would feel.
#include <stdint.h>
#include <string.h>
#include <deque>
#include <string>
#include <vector>
#include "../../include/sse_gateway.h"

namespace {

constexpr size_t CHAN_CAP = 100;

struct Stream {
    bool open = false;          // slot in use
    bool upstream_eof = false;  // response.Body returned EOF / error
    bool closed = false;        // channel closed (after the last line was delivered)
    uint8_t mode = 0;
    std::string pending;        // bytes read from upstream, not yet handed to the GPU
    std::deque<std::string> chan;
    bool terminated = false;    // mode R: agent loop left (finish_reason stop/tool_calls)
    sse_agent_fold *agent = nullptr;
    bool done_sent = false;
};

} // namespace

struct ssegw {
    sse_ctx *ctx = nullptr;
    sse_config cfg{};
    std::vector<Stream> streams;
    std::string last_error;
};

extern "C" {

ssegw *ssegw_new(int device, uint32_t max_conns, uint32_t bytes_per_batch, int *status) {
    ssegw *g = new ssegw();
    sse_default_config(&g->cfg, max_conns, bytes_per_batch);
    g->cfg.carry_slot_bytes = 65536;
    int rc = sse_init(device, &g->cfg, &g->ctx);
    if (status) *status = rc;
    if (rc != SSE_OK) { delete g; return nullptr; }   // no CUDA device: error, never a CPU path
    g->streams.resize(max_conns);
    return g;
}

void ssegw_free(ssegw *g) {
    if (!g) return;
    for (auto &s : g->streams) if (s.agent) sse_agent_free(s.agent);
    sse_destroy(g->ctx);
    delete g;
}

// StreamChatCompletions: allocates a connection slot and its channel. Returns the stream id or -1 (no free slot).
int ssegw_stream_chat_completions(ssegw *g, uint8_t mode) {
    for (size_t i = 0; i < g->streams.size(); i++) {
        Stream &s = g->streams[i];
        if (!s.open) {
            if (s.agent) sse_agent_reset(s.agent); else s.agent = sse_agent_new();
            s = Stream{ true, false, false, mode, {}, {}, false, s.agent, false };
            if (sse_reset_conn(g->ctx, (uint32_t)i) != SSE_OK) return -1;
            return (int)i;
        }
    }
    return -1;
}

// Bytes from response.Body. Returns the number of bytes accepted (0 when the channel is full: back-pressure).
size_t ssegw_upstream_write(ssegw *g, int id, const uint8_t *data, size_t n) {
    Stream &s = g->streams[(size_t)id];
    if (!s.open || s.upstream_eof || s.closed) return 0;
    if (s.chan.size() >= CHAN_CAP) return 0;
    s.pending.append((const char *)data, n);
    return n;
}

void ssegw_upstream_close(ssegw *g, int id) { g->streams[(size_t)id].upstream_eof = true; }

// One batcher tick. Returns the number of frames delivered, or a negative sse_status.
int ssegw_pump(ssegw *g) {
    int slot; sse_batch b;
    int rc = sse_acquire(g->ctx, &slot, &b);
    if (rc != SSE_OK) return rc;
    std::vector<uint32_t> who;
    uint32_t off = 0, n = 0;
    for (size_t i = 0; i < g->streams.size(); i++) {
        Stream &s = g->streams[i];
        if (!s.open || s.closed || s.pending.empty()) continue;
        if (s.chan.size() >= CHAN_CAP) continue;                      // receiver is slow: leave the bytes queued
        size_t take = s.pending.size();
        if (n >= b.max_segs || off + take + 16 > b.in_arena_bytes) break;
        memcpy(b.in_arena + off, s.pending.data(), take);
        b.segs[n] = sse_seg{ (uint32_t)i, off, (uint32_t)take, s.mode, 0, 0 };
        off = (off + (uint32_t)take + 15u) & ~15u;
        s.pending.clear();
        who.push_back((uint32_t)i);
        n++;
    }
    int delivered = 0;
    if (n) {
        rc = sse_submit(g->ctx, slot, n, off);
        sse_result res;
        if (rc == SSE_OK) rc = sse_collect(g->ctx, slot, &res);
        if (rc != SSE_OK) {
            // a failed batch fails its streams: the client sees EOF, exactly like an upstream read error (provider.go:323-330)
            for (uint32_t i : who) g->streams[i].closed = true;
            sse_release(g->ctx, slot);
            return rc;
        }
        for (uint32_t k = 0; k < n; k++) {
            Stream &s = g->streams[who[k]];
            const sse_run *run = &res.segs[k].run;
            for (;;) {
                for (uint32_t f = run->frame_first; f < run->frame_first + run->frame_count; f++) {
                    s.chan.emplace_back((const char *)sse_at(&res, res.frames[f].off), res.frames[f].len);   // fresh copy per element
                    delivered++;
                }
                if (run->next == SSE_NONE) break;
                run = &res.runs[run->next];
            }
            if (s.mode & SSE_MODE_R) sse_agent_feed(s.agent, &res, k);
            if (res.segs[k].flags & SSE_SEG_TERMINATED) s.terminated = true;
            if (res.segs[k].flags & SSE_SEG_DEAD) s.closed = true;
        }
    }
    sse_release(g->ctx, slot);
    for (auto &s : g->streams)
        if (s.open && s.upstream_eof && s.pending.empty()) s.closed = true;   // tail (if any) stays on the device and is dropped
    return delivered;
}

// `line, ok := <-streamCh`: 1 = element copied to buf (*n bytes), 0 = nothing available yet, -1 = channel closed and drained.
int ssegw_recv(ssegw *g, int id, uint8_t *buf, size_t cap, size_t *n) {
    Stream &s = g->streams[(size_t)id];
    if (!s.chan.empty()) {
        const std::string &e = s.chan.front();
        *n = e.size();
        if (e.size() > cap) return SSE_ERR_ARG;
        memcpy(buf, e.data(), e.size());
        s.chan.pop_front();
        return 1;
    }
    return s.closed ? -1 : 0;
}

void ssegw_release_stream(ssegw *g, int id) { g->streams[(size_t)id].open = false; }

// ---- mcp.Agent.RunWithStream view of a mode-R stream (one iteration)
// Next element of the middleware channel: forwarded frames, then exactly one "data: [DONE]\n\n" once the iteration is
// over (finish_reason stop/tool_calls seen, or the provider channel closed). Same return codes as ssegw_recv.
int ssegw_agent_recv(ssegw *g, int id, uint8_t *buf, size_t cap, size_t *n) {
    Stream &s = g->streams[(size_t)id];
    int rc = ssegw_recv(g, id, buf, cap, n);
    if (rc == 1) return 1;
    const bool over = s.terminated || rc == -1;
    if (over && !s.done_sent) {
        static const char done[] = "data: [DONE]\n\n";
        *n = sizeof(done) - 1;
        if (*n > cap) return SSE_ERR_ARG;
        memcpy(buf, done, *n);
        s.done_sent = true;
        return 1;
    }
    return over ? -1 : 0;
}
sse_bytes ssegw_agent_content(ssegw *g, int id) { return sse_agent_content(g->streams[(size_t)id].agent); }
int ssegw_agent_has_tool_calls(ssegw *g, int id) { return sse_agent_has_tool_calls(g->streams[(size_t)id].agent); }
int ssegw_agent_terminated(ssegw *g, int id, int *finish) { return sse_agent_terminated(g->streams[(size_t)id].agent, finish); }
size_t ssegw_agent_tool_calls(ssegw *g, int id, sse_tool_call *calls, size_t cap) {
    return sse_agent_tool_calls(g->streams[(size_t)id].agent, calls, cap);
}

} // extern "C"
