// Package ssegpu binds libssegpu.so (include/sse_gpu.h) through cgo.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain (SURVEY.md section 8c). This file is the
// binding a maintainer adds to the gateway; it is written against the C ABI only and keeps no Go pointer on the C
// side (all buffers are cudaHostAlloc'd by the library and viewed with unsafe.Slice).
package ssegpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../inference_gateway_b200 -lssegpu -Wl,-rpath,${SRCDIR}/../../inference_gateway_b200
#include "sse_gpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

const (
	ModeP     = uint8(C.SSE_MODE_P)
	ModeR     = uint8(C.SSE_MODE_R)
	ModeParse = uint8(C.SSE_MODE_PARSE)
)

// Context is one GPU. It must be driven from ONE goroutine locked to its OS thread (the per-GPU batcher).
type Context struct{ c *C.sse_ctx }

func statusErr(where string, rc C.int) error {
	if rc == C.SSE_OK {
		return nil
	}
	msg := C.GoString(C.sse_strerror(rc))
	if rc == C.SSE_ERR_CUDA {
		msg += ": " + C.GoString(C.sse_last_cuda_error())
	}
	return fmt.Errorf("%s: %s (%d)", where, msg, int(rc))
}

// NewContext fails when no CUDA device is present: there is no CPU fallback.
func NewContext(device int, maxConns, bytesPerBatch uint32) (*Context, error) {
	runtime.LockOSThread()
	var cfg C.sse_config
	C.sse_default_config(&cfg, C.uint32_t(maxConns), C.uint32_t(bytesPerBatch))
	var ctx *C.sse_ctx
	if err := statusErr("sse_init", C.sse_init(C.int(device), &cfg, &ctx)); err != nil {
		return nil, err
	}
	return &Context{c: ctx}, nil
}

func (x *Context) Close() { C.sse_destroy(x.c) }

// Batch is one acquired slot: Arena and Segs alias pinned memory owned by the library.
type Batch struct {
	Slot  C.int
	Arena []byte
	Segs  []C.sse_seg
	nSegs uint32
	off   uint32
}

func (x *Context) Acquire() (*Batch, error) {
	var slot C.int
	var b C.sse_batch
	if err := statusErr("sse_acquire", C.sse_acquire(x.c, &slot, &b)); err != nil {
		return nil, err
	}
	return &Batch{Slot: slot,
		Arena: unsafe.Slice((*byte)(unsafe.Pointer(b.in_arena)), int(b.in_arena_bytes)),
		Segs:  unsafe.Slice(b.segs, int(b.max_segs))}, nil
}

// Add copies the bytes read from one connection since the previous batch (at most one segment per connection).
func (b *Batch) Add(conn uint32, mode uint8, data []byte) error {
	if int(b.nSegs) >= len(b.Segs) || int(b.off)+len(data)+16 > len(b.Arena) {
		return errors.New("ssegpu: batch full")
	}
	copy(b.Arena[b.off:], data)
	b.Segs[b.nSegs] = C.sse_seg{conn: C.uint32_t(conn), in_off: C.uint32_t(b.off), in_len: C.uint32_t(len(data)), mode: C.uint8_t(mode)}
	b.nSegs++
	b.off = (b.off + uint32(len(data)) + 15) &^ 15
	return nil
}

func (x *Context) Submit(b *Batch) error {
	return statusErr("sse_submit", C.sse_submit(x.c, b.Slot, C.uint32_t(b.nSegs), C.uint32_t(b.off)))
}

// Result views the pinned result buffers; valid until Release.
type Result struct{ r C.sse_result }

func (x *Context) Collect(b *Batch) (*Result, error) {
	res := &Result{}
	if err := statusErr("sse_collect", C.sse_collect(x.c, b.Slot, &res.r)); err != nil {
		return nil, err
	}
	return res, nil
}

func (x *Context) Release(b *Batch) error { return statusErr("sse_release", C.sse_release(x.c, b.Slot)) }
func (x *Context) ResetConn(conn uint32) error {
	return statusErr("sse_reset_conn", C.sse_reset_conn(x.c, C.uint32_t(conn)))
}

// Frames calls fn with a FRESH copy of every frame of segment i, in order: the receiver of the reference's
// chan []byte owns each element (provider.go:322), so the pinned arena must not be handed out.
func (r *Result) Frames(i uint32, fn func(line []byte)) (terminated bool) {
	segs := unsafe.Slice(r.r.segs, int(r.r.n_segs))
	frames := unsafe.Slice(r.r.frames, int(r.r.n_frames))
	runs := unsafe.Slice(r.r.runs, int(r.r.n_runs))
	run := segs[i].run
	for {
		for k := run.frame_first; k < run.frame_first+run.frame_count; k++ {
			f := frames[k]
			// arena offset: >= in_base is a span of the batch's own input arena (zero-copy frame), else the out arena
			src := unsafe.Slice((*byte)(unsafe.Pointer(C.sse_at(&r.r, f.off))), int(f.len))
			line := make([]byte, int(f.len))
			copy(line, src)
			fn(line)
		}
		if run.next == C.SSE_NONE {
			break
		}
		run = runs[run.next]
	}
	return segs[i].flags&C.SSE_SEG_TERMINATED != 0
}
