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
constexpr size_t PENDING_CAP = 4u << 20;   // per-stream bytes waiting for a batch

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
    size_t rr = 0;              // the batcher starts its scan one stream further every tick (no slot is favoured)
};

extern "C" {

ssegw *ssegw_new(int device, uint32_t max_conns, uint32_t bytes_per_batch, int *status) {
    ssegw *g = new ssegw();
    // worst-case result capacities: whatever the upstreams send (a flood of one-byte lines, of empty tool-call elements ...), a
    // batch cannot overflow, so one connection can never fail the batch of the others
    sse_worst_case_config(&g->cfg, max_conns, bytes_per_batch);
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
    // bytes not yet handed to the GPU are bounded like a socket buffer: a reader that outruns the batcher is told to wait
    const size_t room = s.pending.size() < PENDING_CAP ? PENDING_CAP - s.pending.size() : 0;
    if (n > room) n = room;
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
    const size_t ns = g->streams.size();
    const size_t start = ns ? g->rr++ % ns : 0;
    for (size_t q = 0; q < ns && n < b.max_segs; q++) {
        const size_t i = (start + q) % ns;
        Stream &s = g->streams[i];
        if (!s.open || s.closed || s.pending.empty()) continue;
        if (s.chan.size() >= CHAN_CAP) continue;                      // receiver is slow: leave the bytes queued
        // a stream takes what fits: the rest stays queued for the next tick (the carry state on the device makes any cut legal)
        const size_t room = b.in_arena_bytes > off + 16u ? (size_t)(b.in_arena_bytes - off - 16u) : 0;
        const size_t take = s.pending.size() < room ? s.pending.size() : room;
        if (take == 0) continue;
        memcpy(b.in_arena + off, s.pending.data(), take);
        b.segs[n] = sse_seg{ (uint32_t)i, off, (uint32_t)take, s.mode, 0, 0 };
        off = (off + (uint32_t)take + 15u) & ~15u;
        s.pending.erase(0, take);
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
            // a record the library could not decode (SSE_ERR_UNDECODED): the accumulators would be silently short, so the stream is failed
            if ((s.mode & SSE_MODE_R) && sse_agent_feed(s.agent, &res, k) != SSE_OK) s.closed = true;
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

// ---- handleStreamingRequest (api/routes.go:129-232): the raw proxy loop reads and writes in one goroutine
int ssegw_proxy_stream(ssegw *g) { return ssegw_stream_chat_completions(g, SSE_MODE_P); }
int ssegw_proxy_step(ssegw *g, int id, uint8_t *buf, size_t cap, size_t *n) {
    for (;;) {
        const int rc = ssegw_recv(g, id, buf, cap, n);
        if (rc != 1) return rc;              // 0: nothing yet; -1: ReadBytes error / EOF, tail dropped (:187-195)
        if (*n == 0) continue;               // :197-199
        return 1;                            // :220-228 write + flush
    }
}

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

// ---- handleMCPStreamingRequest's writer, one channel element (api/middlewares/mcp.go:253-299)
namespace {

// encoding/json's checkValid + a walk of the top-level object, enough to tell whether
// json.Unmarshal(data, &struct{ Error string `json:"error"` }) returns nil. Iterative (Go's nesting limit is 10000).
struct ErrSniff {
    const uint8_t *p, *e;
    bool ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; return p < e; }
    static int hexv(uint8_t c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }
    // scans a string starting at the opening quote; appends the unquoted bytes to *out when given
    bool str(std::string *out) {
        p++;
        while (p < e) {
            uint8_t c = *p;
            if (c == '"') { p++; return true; }
            if (c < 0x20) return false;
            if (c != '\\') { if (out) out->push_back((char)c); p++; continue; }
            if (++p >= e) return false;
            c = *p++;
            switch (c) {
            case '"': case '\\': case '/': if (out) out->push_back((char)c); break;
            case 'b': if (out) out->push_back('\b'); break;
            case 'f': if (out) out->push_back('\f'); break;
            case 'n': if (out) out->push_back('\n'); break;
            case 'r': if (out) out->push_back('\r'); break;
            case 't': if (out) out->push_back('\t'); break;
            case 'u': {
                if (e - p < 4) return false;
                int v = 0;
                for (int k = 0; k < 4; k++) { int h = hexv(p[k]); if (h < 0) return false; v = v * 16 + h; }
                p += 4;
                // only ASCII matters for the comparison with "error"; anything else makes the key differ
                if (out) out->push_back(v < 0x80 ? (char)v : (char)0xFF);
                break;
            }
            default: return false;
            }
        }
        return false;
    }
    bool number() {
        if (p < e && *p == '-') p++;
        if (p >= e) return false;
        if (*p == '0') p++;
        else if (*p >= '1' && *p <= '9') { while (p < e && *p >= '0' && *p <= '9') p++; }
        else return false;
        if (p < e && *p == '.') { p++; if (p >= e || *p < '0' || *p > '9') return false; while (p < e && *p >= '0' && *p <= '9') p++; }
        if (p < e && (*p == 'e' || *p == 'E')) {
            p++;
            if (p < e && (*p == '+' || *p == '-')) p++;
            if (p >= e || *p < '0' || *p > '9') return false;
            while (p < e && *p >= '0' && *p <= '9') p++;
        }
        return true;
    }
    bool lit(const char *w) { size_t n = strlen(w); if ((size_t)(e - p) < n || memcmp(p, w, n) != 0) return false; p += n; return true; }
    static bool is_error_key(const std::string &k) {
        if (k.size() != 5) return false;
        static const char want[] = "error";
        for (int i = 0; i < 5; i++) { char c = k[i]; if (c >= 'A' && c <= 'Z') c = (char)(c + 32); if (c != want[i]) return false; }
        return true;
    }
    // returns true when Unmarshal would return nil
    bool run() {
        std::vector<uint8_t> stk;     // 1 = object, 0 = array
        bool type_ok = true;
        if (!ws() || *p != '{') return false;       // (the caller checked the "data: {" prefix; kept for completeness)
        // state machine over values
        enum { VALUE, AFTER } st = VALUE;
        bool pending_error_key = false;             // the value about to be read belongs to a top-level "error" key
        for (;;) {
            if (st == VALUE) {
                if (!ws()) return false;
                const uint8_t c = *p;
                const bool top = stk.size() == 1 && stk.back() == 1;
                const bool chk = top && pending_error_key;
                if (c == '{' || c == '[') {
                    if (chk) type_ok = false;
                    pending_error_key = false;
                    if (stk.size() >= 10000) return false;
                    stk.push_back(c == '{');
                    p++;
                    if (!ws()) return false;
                    if (c == '{') {
                        if (*p == '}') { p++; stk.pop_back(); st = AFTER; continue; }
                        // key
                        if (*p != '"') return false;
                        std::string key; const bool want_key = stk.size() == 1;
                        if (!str(want_key ? &key : nullptr)) return false;
                        if (want_key) pending_error_key = is_error_key(key);
                        if (!ws() || *p != ':') return false;
                        p++;
                        continue;
                    }
                    if (*p == ']') { p++; stk.pop_back(); st = AFTER; continue; }
                    continue;
                }
                if (stk.empty()) return false;
                if (c == '"') { if (!str(nullptr)) return false; }
                else if (c == '-' || (c >= '0' && c <= '9')) { if (!number()) return false; if (chk) type_ok = false; }
                else if (c == 't') { if (!lit("true")) return false; if (chk) type_ok = false; }
                else if (c == 'f') { if (!lit("false")) return false; if (chk) type_ok = false; }
                else if (c == 'n') { if (!lit("null")) return false; }
                else return false;
                pending_error_key = false;
                st = AFTER;
                continue;
            }
            // AFTER a value
            if (stk.empty()) { ws(); return p == e && type_ok; }
            if (!ws()) return false;
            if (stk.back() == 1) {
                if (*p == '}') { p++; stk.pop_back(); continue; }
                if (*p != ',') return false;
                p++;
                if (!ws() || *p != '"') return false;
                std::string key; const bool want_key = stk.size() == 1;
                if (!str(want_key ? &key : nullptr)) return false;
                if (want_key) pending_error_key = is_error_key(key);
                if (!ws() || *p != ':') return false;
                p++;
                st = VALUE;
            } else {
                if (*p == ']') { p++; stk.pop_back(); continue; }
                if (*p != ',') return false;
                p++;
                st = VALUE;
            }
        }
    }
};

bool contains(const uint8_t *h, size_t n, const char *needle) {
    const size_t m = strlen(needle);
    if (n < m) return false;
    for (size_t i = 0; i + m <= n; i++) if (h[i] == (uint8_t)needle[0] && memcmp(h + i, needle, m) == 0) return true;
    return false;
}

} // namespace

int ssegw_mcp_writer_step(const uint8_t *frame, size_t n, int *set_503) {
    static const char done[] = "data: [DONE]\n\n";
    if (set_503) *set_503 = 0;
    if (n == sizeof(done) - 1 && memcmp(frame, done, n) == 0) return 1;                 // mcp.go:261-268
    if (n >= 7 && memcmp(frame, "data: {", 7) == 0 && contains(frame, n, "\"error\"")) {   // mcp.go:272-280
        ErrSniff s{ frame + 6, frame + n };
        if (s.run() && set_503) *set_503 = 1;
    }
    return 0;
}

} // extern "C"
