// Package baseline: the Go denominator of BASELINE.md (SURVEY.md 8(d), "CPU baseline timed beside it").
//
// NOT compiled in this repository (the image has no Go toolchain). Drop this file into the reference module
// (github.com/inference-gateway/inference-gateway, e.g. under tests/) wherever Go 1.26 exists and run
//
//	python tools/dump_workload.py C4 65536 /tmp/c4.bin          # this repository: the bench's exact bytes
//	SSE_WORKLOAD=/tmp/c4.bin go test -run xxx -bench Stream -benchtime 5x ./tests/
//
// It drives the UNMODIFIED reference code of the path:
//   - mode P: core.ProviderImpl.StreamChatCompletions (providers/core/provider.go:277-344, the ReadBytes('\n') reader
//     goroutine and its chan []byte of capacity 100) consumed the way api/routes.go:600-625 does (write every line);
//   - mode R: the same producer consumed by the statements of mcp/agent.go:169-248 (TrimSpace, Contains "[DONE]",
//     HasPrefix "data: ", Sprintf reframe, json.Unmarshal into types.CreateChatCompletionStreamResponse, content
//     accumulation through the MessageContent union, early termination). agentImpl.RunWithStream itself needs a live MCP
//     client, so its inner loop is restated here line by line; everything it calls is the reference's own code.
//
// One "chunk" = one emitted frame, as in bench.py. Streams run concurrently on GOMAXPROCS workers.
package baseline

import (
	"bytes"
	"context"
	"crypto/sha256"
	"encoding/binary"
	"encoding/json"
	"fmt"
	"io"
	"net/http"
	"os"
	"runtime"
	"strings"
	"sync"
	"sync/atomic"
	"testing"

	"github.com/inference-gateway/inference-gateway/logger"
	"github.com/inference-gateway/inference-gateway/providers/constants"
	"github.com/inference-gateway/inference-gateway/providers/core"
	"github.com/inference-gateway/inference-gateway/providers/types"
)

// bodyClient answers every request with one prepared upstream body (client.Client, providers/client/client.go:16-20).
type bodyClient struct{ body []byte }

func (c *bodyClient) Do(*http.Request) (*http.Response, error) {
	return &http.Response{StatusCode: http.StatusOK, Body: io.NopCloser(bytes.NewReader(c.body))}, nil
}
func (c *bodyClient) Get(string) (*http.Response, error)                  { return c.Do(nil) }
func (c *bodyClient) Post(string, string, string) (*http.Response, error) { return c.Do(nil) }

func loadWorkload(tb testing.TB) [][]byte {
	path := os.Getenv("SSE_WORKLOAD")
	if path == "" {
		tb.Skip("SSE_WORKLOAD not set (tools/dump_workload.py writes it)")
	}
	raw, err := os.ReadFile(path)
	if err != nil {
		tb.Fatal(err)
	}
	var streams [][]byte
	for off := 0; off+4 <= len(raw); {
		n := int(binary.LittleEndian.Uint32(raw[off:]))
		off += 4
		streams = append(streams, raw[off:off+n])
		off += n
	}
	return streams
}

func newProvider(body []byte) *core.ProviderImpl {
	id := constants.OllamaID
	return &core.ProviderImpl{ID: &id, Name: "bench", Endpoints: types.Endpoints{Chat: "/v1/chat/completions"},
		Client: &bodyClient{body: body}, Logger: logger.NewNoopLogger()}
}

// passthrough: api/routes.go:600-625 without gin (Write + Flush per line are replaced by io.Discard).
func passthrough(ctx context.Context, body []byte) (frames int64) {
	ch, err := newProvider(body).StreamChatCompletions(ctx, types.CreateChatCompletionRequest{})
	if err != nil {
		panic(err)
	}
	for line := range ch {
		_, _ = io.Discard.Write(line)
		frames++
	}
	return frames
}

// reframe: mcp/agent.go:169-248, one iteration.
func reframe(ctx context.Context, body []byte, sink chan<- []byte) (frames int64) {
	ch, err := newProvider(body).StreamChatCompletions(ctx, types.CreateChatCompletionRequest{})
	if err != nil {
		panic(err)
	}
	var responseBodyBuilder strings.Builder
	assistantMessage := types.Message{Role: types.Assistant}
	streamComplete := false
	for !streamComplete {
		line, ok := <-ch
		if !ok {
			break
		}
		trimmed := strings.TrimSpace(string(line))
		if strings.Contains(trimmed, "[DONE]") {
			continue
		}
		if !strings.HasPrefix(trimmed, "data: ") {
			continue
		}
		payload := strings.TrimPrefix(trimmed, "data: ")
		if payload == "" {
			continue
		}
		frame := fmt.Sprintf("data: %s\n\n", payload)
		sink <- []byte(frame)
		responseBodyBuilder.WriteString(frame)
		frames++
		var chunk types.CreateChatCompletionStreamResponse
		if err := json.Unmarshal([]byte(payload), &chunk); err != nil {
			continue
		}
		if len(chunk.Choices) == 0 {
			continue
		}
		choice := chunk.Choices[0]
		if choice.Delta.Content != "" { // agent.go:211-222: the O(n^2) accumulation through the MessageContent union
			if cur, err := assistantMessage.Content.AsMessageContent0(); err == nil {
				_ = assistantMessage.Content.FromMessageContent0(cur + choice.Delta.Content)
			} else {
				_ = assistantMessage.Content.FromMessageContent0(choice.Delta.Content)
			}
		}
		if fr := string(choice.FinishReason); fr == "stop" || fr == "tool_calls" {
			streamComplete = true
		}
	}
	go func() { // the reader goroutine of an abandoned stream blocks on its channel: drain it as ctx cancellation would
		for range ch {
		}
	}()
	return frames
}

func run(b *testing.B, mode string) {
	streams := loadWorkload(b)
	workers := runtime.GOMAXPROCS(0)
	ctx := context.Background()
	sink := make(chan []byte, 4096)
	go func() {
		for range sink {
		}
	}()
	var total int64
	b.ResetTimer()
	for i := 0; i < b.N; i++ {
		var next int64 = -1
		var wg sync.WaitGroup
		for w := 0; w < workers; w++ {
			wg.Add(1)
			go func() {
				defer wg.Done()
				for {
					k := atomic.AddInt64(&next, 1)
					if int(k) >= len(streams) {
						return
					}
					if mode == "P" {
						atomic.AddInt64(&total, passthrough(ctx, streams[k]))
					} else {
						atomic.AddInt64(&total, reframe(ctx, streams[k], sink))
					}
				}
			}()
		}
		wg.Wait()
	}
	b.StopTimer()
	b.ReportMetric(float64(total)/b.Elapsed().Seconds(), "chunks/s")
	b.ReportMetric(float64(workers), "cores")
}

// TestDumpParity writes, per stream, the SHA-256 of what the unmodified reference emits in mode P and in mode R
// (frames concatenated). `python tools/check_go_dump.py $SSE_WORKLOAD $SSE_DUMP` compares it with this repository's oracle:
// the byte-level pin that DESIGN.md section 6 says is missing ("parity unpinned").
//
//	SSE_WORKLOAD=/tmp/c4.bin SSE_DUMP=/tmp/c4_go.jsonl go test -run TestDumpParity ./tests/
func TestDumpParity(t *testing.T) {
	streams := loadWorkload(t)
	path := os.Getenv("SSE_DUMP")
	if path == "" {
		t.Skip("SSE_DUMP not set")
	}
	f, err := os.Create(path)
	if err != nil {
		t.Fatal(err)
	}
	defer f.Close()
	ctx := context.Background()
	for i, body := range streams {
		hp := sha256.New()
		ch, err := newProvider(body).StreamChatCompletions(ctx, types.CreateChatCompletionRequest{})
		if err != nil {
			t.Fatal(err)
		}
		nP := 0
		for line := range ch {
			hp.Write(line)
			nP++
		}
		hr := sha256.New()
		sink := make(chan []byte, 1<<16)
		nR := reframe(ctx, body, sink)
		close(sink)
		for fr := range sink {
			hr.Write(fr)
		}
		fmt.Fprintf(f, "{\"stream\":%d,\"p_frames\":%d,\"p_sha256\":\"%x\",\"r_frames\":%d,\"r_sha256\":\"%x\"}\n", i, nP, hp.Sum(nil), nR, hr.Sum(nil))
	}
}

func BenchmarkStreamPassthrough(b *testing.B) { run(b, "P") }
func BenchmarkStreamReframe(b *testing.B)     { run(b, "R") }
