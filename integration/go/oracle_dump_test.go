// Dumps golden vectors from the REAL reference functions, in the format tests/golden/*.json uses, so that the C
// oracle can be validated wherever Go 1.26 exists:  go test ./integration/go -run TestDumpGolden -v
// NOT RUN IN THIS REPOSITORY (no Go toolchain in the image); until it is, parity is "unpinned" at byte level.
package ssegpu

import (
	"encoding/json"
	"os"
	"strings"
	"testing"

	"github.com/inference-gateway/inference-gateway/providers/types"
)

type goldenChunk struct {
	Payload   string `json:"payload"`
	OK        bool   `json:"json_ok"`
	NChoices  int    `json:"n_choices"`
	Content   string `json:"content"`
	Finish    string `json:"finish_reason"`
	HasUsage  bool   `json:"has_usage"`
	Prompt    int64  `json:"prompt_tokens"`
	ToolCalls int    `json:"tool_calls"`
}

func TestDumpGolden(t *testing.T) {
	in, err := os.ReadFile("../../tests/golden/payloads.txt") // one JSON document per line (tests/corpus.py TRICKY)
	if err != nil {
		t.Skip("no payload list")
	}
	var out []goldenChunk
	for _, line := range strings.Split(strings.TrimRight(string(in), "\n"), "\n") {
		var resp types.CreateChatCompletionStreamResponse
		g := goldenChunk{Payload: line}
		if err := json.Unmarshal([]byte(line), &resp); err == nil {
			g.OK, g.NChoices = true, len(resp.Choices)
			if len(resp.Choices) > 0 {
				g.Content, g.Finish = resp.Choices[0].Delta.Content, string(resp.Choices[0].FinishReason)
				if resp.Choices[0].Delta.ToolCalls != nil {
					g.ToolCalls = len(*resp.Choices[0].Delta.ToolCalls)
				}
			}
			if resp.Usage != nil {
				g.HasUsage, g.Prompt = true, resp.Usage.PromptTokens
			}
		}
		out = append(out, g)
	}
	b, _ := json.MarshalIndent(out, "", " ")
	_ = os.WriteFile("../../tests/golden/go_unmarshal.json", b, 0o644)
}
