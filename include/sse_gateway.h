/*
 * sse_gateway.h -- host-side mirror of the reference's streaming interfaces, above the C ABI of sse_gpu.h
 * (implemented in inference_gateway_b200/csrc/sse_gateway.cpp, exported by libssegpu.so).
 *
 *   ssegw_stream_chat_completions   core.IProvider.StreamChatCompletions (providers/core/interfaces.go:22; provider.go:277-344)
 *   ssegw_upstream_write / _close   response.Body bytes / EOF-or-error (provider.go:322-330)
 *   ssegw_pump                      one tick of the per-GPU batcher
 *   ssegw_recv                      `line, ok := <-streamCh` (api/routes.go:602-606, mcp/agent.go:171)
 *   ssegw_agent_recv and friends    mcp.Agent.RunWithStream for one iteration (mcp/agent.go:126-290, final [DONE] :140-143)
 *   ssegw_proxy_stream / _step    handleStreamingRequest, the raw /proxy/:provider/... stream loop (api/routes.go:129-232)
 *   ssegw_mcp_writer_step           one turn of handleMCPStreamingRequest's writer (api/middlewares/mcp.go:253-299): the
 *                                   terminal-frame rule (:261-268) and the upstream-error sniff that may set 503 (:272-280)
 */
#ifndef SSE_GATEWAY_H
#define SSE_GATEWAY_H
#include "sse_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssegw ssegw;

ssegw *ssegw_new(int device, uint32_t max_conns, uint32_t bytes_per_batch, int *status);
void   ssegw_free(ssegw *g);
int    ssegw_stream_chat_completions(ssegw *g, uint8_t mode);
size_t ssegw_upstream_write(ssegw *g, int stream, const uint8_t *data, size_t n);
void   ssegw_upstream_close(ssegw *g, int stream);
int    ssegw_pump(ssegw *g);
int    ssegw_recv(ssegw *g, int stream, uint8_t *buf, size_t cap, size_t *n);
void   ssegw_release_stream(ssegw *g, int stream);
int    ssegw_agent_recv(ssegw *g, int stream, uint8_t *buf, size_t cap, size_t *n);
sse_bytes ssegw_agent_content(ssegw *g, int stream);
int    ssegw_agent_has_tool_calls(ssegw *g, int stream);
int    ssegw_agent_terminated(ssegw *g, int stream, int *finish);
size_t ssegw_agent_tool_calls(ssegw *g, int stream, sse_tool_call *calls, size_t cap);

/* The raw proxy stream handler (api/routes.go:129-232, taken when Accept is exactly text/event-stream, :118): reader and
 * writer in one loop, no channel in between. ssegw_proxy_stream opens the upstream body as a mode P stream (fed through
 * ssegw_upstream_write / _close like any other); ssegw_proxy_step is one turn of the c.Stream callback (:178-231):
 *   1  a line (including its '\n') was copied to buf: the handler writes and flushes it (:220-228) and goes on;
 *   0  no complete line has arrived yet;
 *  -1  ReadBytes returned an error (EOF included): the callback returns false and the unterminated tail is never
 *      written (:187-195).
 * A zero-length element is skipped (:197-199; unreachable, ReadBytes never returns an empty slice with a nil error). */
int    ssegw_proxy_stream(ssegw *g);
int    ssegw_proxy_step(ssegw *g, int stream, uint8_t *buf, size_t cap, size_t *n);

/* The MCP writer (api/middlewares/mcp.go:253-299) for one element of the agent's channel. The frame is always written
 * unchanged. Returns 1 when the stream ends after this write (the frame is byte-equal to "data: [DONE]\n\n", :261-268),
 * else 0. *set_503 becomes 1 when the reference would call WriteHeader(503) for this frame (:272-280): it starts with
 * "data: {", contains "\"error\"", and json.Unmarshal(frame[6:], &struct{ Error string `json:"error"` }) returns nil --
 * i.e. the rest is valid JSON whose top-level keys matching "error" (exactly or case-insensitively) all hold a string or
 * null. Pure host code (no device involved): such frames are rare and the status is not part of the byte stream. */
int ssegw_mcp_writer_step(const uint8_t *frame, size_t n, int *set_503);

#ifdef __cplusplus
}
#endif
#endif
