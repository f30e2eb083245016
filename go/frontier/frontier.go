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

// SchedTables are the limits and the running-StepRun counts the batch does not hold (dag.go:1780-1961).
type SchedTables struct {
	StoryLimit       []int32  // Story.spec.policy.concurrency per story key
	StoryRunningBase []uint32 // nil = zeros
	QueueLimit       []int32  // scheduling.queues[q].concurrency per queue key
	QueueAgingS      []int32  // scheduling.queues[q].priorityAgingSeconds
	QueueRunningBase []uint32 // nil = zeros
	GlobalLimit      int32    // scheduling.globalConcurrency
	GlobalBase       uint32
}

// SchedResult holds the schedule records (BF_SCHED_STRIDE(words) bytes per run: bf_sched_header, launch,
// queued_story, queued_sched masks) and the totals the "(%d running, limit %d)" messages are formatted from.
type SchedResult struct {
	Records       []byte
	StoryRunning  []uint32
	QueueRunning  []uint32
	QueueMaxPrio  []int32
	GlobalRunning uint32
}

// Schedule applies enforceStoryConcurrency / enforceSchedulingLimits to the ready sets of the batch Eval just
// evaluated: it replaces the three cluster-wide LISTs per reconcile (dag.go:1863-1920) by one reduction per tick.
func (c *Ctx) Schedule(b *Batch, runs []C.bf_sched_run, t *SchedTables) (*SchedResult, error) {
	words := uint32(b.Layout.words)
	stride := (16 + 12*words + 15) &^ 15
	res := &SchedResult{
		Records:      make([]byte, int(b.N)*int(stride)),
		StoryRunning: make([]uint32, len(t.StoryLimit)),
		QueueRunning: make([]uint32, len(t.QueueLimit)),
		QueueMaxPrio: make([]int32, len(t.QueueLimit)),
	}
	ct := C.bf_sched_tables{
		struct_size:         C.uint32_t(unsafe.Sizeof(C.bf_sched_tables{})),
		n_stories:           C.uint32_t(len(t.StoryLimit)),
		n_queues:            C.uint32_t(len(t.QueueLimit)),
		global_limit:        C.int32_t(t.GlobalLimit),
		global_running_base: C.uint32_t(t.GlobalBase),
	}
	if len(t.StoryLimit) > 0 {
		ct.story_limit = (*C.int32_t)(unsafe.Pointer(&t.StoryLimit[0]))
	}
	if len(t.StoryRunningBase) > 0 {
		ct.story_running_base = (*C.uint32_t)(unsafe.Pointer(&t.StoryRunningBase[0]))
	}
	if len(t.QueueLimit) > 0 {
		ct.queue_limit = (*C.int32_t)(unsafe.Pointer(&t.QueueLimit[0]))
		ct.queue_aging_s = (*C.int32_t)(unsafe.Pointer(&t.QueueAgingS[0]))
	}
	if len(t.QueueRunningBase) > 0 {
		ct.queue_running_base = (*C.uint32_t)(unsafe.Pointer(&t.QueueRunningBase[0]))
	}
	out := C.bf_sched_out{struct_size: C.uint32_t(unsafe.Sizeof(C.bf_sched_out{})), global_running: (*C.uint32_t)(unsafe.Pointer(&res.GlobalRunning))}
	if len(res.Records) > 0 {
		out.records = unsafe.Pointer(&res.Records[0])
	}
	if len(res.StoryRunning) > 0 {
		out.story_running = (*C.uint32_t)(unsafe.Pointer(&res.StoryRunning[0]))
	}
	if len(res.QueueRunning) > 0 {
		out.queue_running = (*C.uint32_t)(unsafe.Pointer(&res.QueueRunning[0]))
		out.queue_max_priority = (*C.int32_t)(unsafe.Pointer(&res.QueueMaxPrio[0]))
	}
	cb := C.bf_batch{struct_size: C.uint32_t(unsafe.Sizeof(C.bf_batch{})), n_runs: C.uint32_t(b.N), layout: b.Layout}
	var rp *C.bf_sched_run
	if len(runs) > 0 {
		rp = &runs[0]
	}
	return res, c.err(C.bf_schedule(c.p, &cb, rp, &ct, &out))
}

// RedriveClosure returns, for each (slot, step), the bit mask of the steps a redrive from that step resets
// (resolveRedriveFromStepSet, storyrun_controller.go:535-558).
func (c *Ctx) RedriveClosure(slots, steps []uint32, words uint32) ([]uint32, error) {
	masks := make([]uint32, len(slots)*int(words))
	if len(slots) == 0 {
		return masks, nil
	}
	return masks, c.err(C.bf_topology_closure(c.p, (*C.uint32_t)(unsafe.Pointer(&slots[0])), (*C.uint32_t)(unsafe.Pointer(&steps[0])),
		C.uint32_t(len(slots)), C.uint32_t(words), (*C.uint32_t)(unsafe.Pointer(&masks[0]))))
}

// Resident is a device-resident batch (row f2): full records travel once, afterwards only deltas.
type Resident struct {
	c      *Ctx
	handle C.uint32_t
	layout C.bf_layout
}

func (c *Ctx) NewResident(layout C.bf_layout, capacity uint32) (*Resident, error) {
	r := &Resident{c: c, layout: layout}
	return r, c.err(C.bf_resident_create(c.p, &layout, C.uint32_t(capacity), &r.handle))
}

func (r *Resident) Close() error { return r.c.err(C.bf_resident_destroy(r.c.p, r.handle)) }

// Upload sends full state records for runs [first, first+n) (new StoryRuns, or a resync).
func (r *Resident) Upload(first, n uint32, records unsafe.Pointer) error {
	return r.c.err(C.bf_resident_upload(r.c.p, r.handle, C.uint32_t(first), C.uint32_t(n), records))
}

// Apply sends the tick's coalesced deltas: what syncStateFromStepRuns (dag.go:965-1009) changed.
func (r *Resident) Apply(deltas []C.bf_delta) error {
	if len(deltas) == 0 {
		return nil
	}
	return r.c.err(C.bf_resident_apply(r.c.p, r.handle, &deltas[0], C.uint32_t(len(deltas))))
}

// Eval runs one pass over runs [0, n) of the resident state and reads the result records back.
func (r *Resident) Eval(n, flags uint32, result unsafe.Pointer) (C.bf_counts, error) {
	var counts C.bf_counts
	return counts, r.c.err(C.bf_resident_eval(r.c.p, r.handle, C.uint32_t(n), C.uint32_t(flags), 0, result, &counts))
}
