// Package frontier is the cgo shim over include/bobrafrontier.h.
//
// SOURCE ONLY: this image has no Go toolchain, so this file is not compiled or tested here.  It shows the
// reference-side binding a bobrapet maintainer would add; the C ABI it binds is exercised by the Python
// ctypes binding (bobrapet_b200/_abi.py) in tests/ and bench.py.
//
// Build (on a box with Go + the built library):
//   CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/bobrapet_b200/lib -lbobrafrontier" go build ./...
package frontier

/*
#include <stdlib.h>
#include "bobrafrontier.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// Ctx owns one GPU.  bf_ctx is thread-safe (internal mutex); the recommended shape is still one batcher
// goroutine per Ctx (see INTEGRATION.md).
type Ctx struct{ p *C.bf_ctx }

type Error struct {
	Status int
	Msg    string
}

func (e *Error) Error() string {
	return fmt.Sprintf("bobrafrontier: %s (%d): %s", C.GoString(C.bf_strerror(C.int(e.Status))), e.Status, e.Msg)
}

func (c *Ctx) err(rc C.int) error {
	if rc == C.BF_OK {
		return nil
	}
	return &Error{Status: int(rc), Msg: C.GoString(C.bf_last_error(c.p))}
}

func New(device int) (*Ctx, error) {
	cfg := C.bf_config{struct_size: C.uint32_t(unsafe.Sizeof(C.bf_config{})), device: C.int32_t(device)}
	var p *C.bf_ctx
	if rc := C.bf_create(&p, &cfg); rc != C.BF_OK {
		return nil, &Error{Status: int(rc), Msg: "bf_create"}
	}
	return &Ctx{p: p}, nil
}

func (c *Ctx) Close() { C.bf_destroy(c.p); c.p = nil }

// Topology is one Story generation packed by the host: CSR over allStorySteps (dag.go:3270) + step flags.
type Topology struct {
	RowPtr    []uint32 // len S+1
	ColIdx    []uint16 // len E
	StepFlags []uint8  // len S, BF_SF_*
	Parallel  []C.bf_parallel_desc
	AllowBits []uint8
}

// PutTopology uploads one topology and returns its slot.  Slices hold no Go pointers, so passing their
// backing arrays for the duration of the call is cgo-legal; the library copies what it keeps.
func (c *Ctx) PutTopology(t *Topology) (uint32, error) {
	ct := C.bf_topology{
		n_steps:    C.uint32_t(len(t.StepFlags)),
		n_edges:    C.uint32_t(len(t.ColIdx)),
		row_ptr:    (*C.uint32_t)(unsafe.Pointer(&t.RowPtr[0])),
		step_flags: (*C.uint8_t)(unsafe.Pointer(&t.StepFlags[0])),
		n_parallel: C.uint32_t(len(t.Parallel)),
	}
	if len(t.ColIdx) > 0 {
		ct.col_idx = (*C.uint16_t)(unsafe.Pointer(&t.ColIdx[0]))
	}
	if len(t.Parallel) > 0 {
		ct.parallel = &t.Parallel[0]
	}
	if len(t.AllowBits) > 0 {
		ct.branch_allow_bits = (*C.uint8_t)(unsafe.Pointer(&t.AllowBits[0]))
		ct.n_branch_allow_bits = C.uint32_t(8 * len(t.AllowBits))
	}
	var slot C.uint32_t
	if err := c.err(C.bf_topology_put(c.p, &ct, &slot)); err != nil {
		return 0, err
	}
	return uint32(slot), nil
}

func (c *Ctx) DropTopology(slot uint32) error { return c.err(C.bf_topology_drop(c.p, C.uint32_t(slot))) }

// Batch is a pair of C-allocated pinned buffers (bf_alloc_pinned) laid out by bf_layout_init; the batcher
// writes state records into State and reads result records from Result.
type Batch struct {
	Layout C.bf_layout
	N      uint32
	State  unsafe.Pointer
	Result unsafe.Pointer
}

func (c *Ctx) NewBatch(stepsMax, childNibbles, fields uint32, capRuns uint32) (*Batch, error) {
	b := &Batch{}
	if rc := C.bf_layout_init(&b.Layout, C.uint32_t(stepsMax), C.uint32_t(childNibbles), C.uint32_t(fields)); rc != C.BF_OK {
		return nil, &Error{Status: int(rc), Msg: "bf_layout_init"}
	}
	if err := c.err(C.bf_alloc_pinned(c.p, C.size_t(capRuns)*C.size_t(b.Layout.state_stride), &b.State)); err != nil {
		return nil, err
	}
	if err := c.err(C.bf_alloc_pinned(c.p, C.size_t(capRuns)*C.size_t(b.Layout.result_stride), &b.Result)); err != nil {
		return nil, err
	}
	return b, nil
}

// Eval runs one frontier pass over the batch (H2D state -> kernels -> D2H results), synchronously and
// bounded; a non-nil error makes the reconciler requeue with backoff or fall back to the Go path.
func (c *Ctx) Eval(b *Batch, flags uint32) (C.bf_counts, error) {
	var counts C.bf_counts
	cb := C.bf_batch{
		struct_size: C.uint32_t(unsafe.Sizeof(C.bf_batch{})),
		n_runs:      C.uint32_t(b.N),
		flags:       C.uint32_t(flags),
		layout:      b.Layout,
		state:       b.State,
		result:      b.Result,
		counts:      &counts,
	}
	return counts, c.err(C.bf_eval(c.p, &cb))
}
