// Package frontier is the cgo binding of include/bobrafrontier.h: the per-reconcile "which steps are ready now?"
// computation of internal/controller/runs/dag.go evaluated for a whole batch of StoryRuns on B200 GPUs.
//
// SOURCE ONLY in this repository: the build image has no Go toolchain.  What keeps it honest without one:
//   - every C call goes through go/frontier/shim.h — flat arguments, structs built on the C stack — and
//     tests/c_abi_harness.c compiles that same header with gcc and performs this file's call sequence on a GPU;
//   - no call passes a pointer to Go memory that holds Go pointers (the cgo rule the previous draft broke): Go only
//     hands over &slice[0] of pointer-free slices, or C / pinned memory it got from the library.
//
// Build (on a box with Go + the built library):
//
//	CGO_CFLAGS="-I${REPO}/include -I${REPO}/go/frontier" CGO_LDFLAGS="-L${REPO}/bobrapet_b200/lib -lbobrafrontier" go build ./...
package frontier

/*
#include <stdlib.h>
#include "shim.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// Ctx owns one GPU.  bf_ctx is thread-safe (internal mutex); the recommended shape is one batcher goroutine
// per Ctx (INTEGRATION.md).
type Ctx struct{ p *C.bf_ctx }

// Error carries a negative bf_status and the library's text.
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

// New creates the context of one device.
func New(device int) (*Ctx, error) {
	var p *C.bf_ctx
	if rc := C.bfgo_create(C.int32_t(device), &p); rc != C.BF_OK {
		return nil, &Error{Status: int(rc), Msg: "bf_create"}
	}
	return &Ctx{p: p}, nil
}

func (c *Ctx) Close() { C.bf_destroy(c.p); c.p = nil }

func u32p(s []uint32) *C.uint32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint32_t)(unsafe.Pointer(&s[0]))
}
func i32p(s []int32) *C.int32_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.int32_t)(unsafe.Pointer(&s[0]))
}
func u8p(s []uint8) *C.uint8_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&s[0]))
}

// ParallelDesc mirrors bf_parallel_desc (8 bytes, no pointers).
type ParallelDesc struct {
	Step       uint16
	Branches   uint16
	AllowFirst uint32
}

// Topology is one Story generation packed by PackStory: CSR over allStorySteps (dag.go:3270) + step flags.
type Topology struct {
	RowPtr    []uint32 // len S+1
	ColIdx    []uint16 // len E
	StepFlags []uint8  // len S, BF_SF_*
	Parallel  []ParallelDesc
	AllowBits []uint8
}

// PutTopology uploads one topology and returns its slot.  BF_ETOPO means what validateRuntimeDependencyGraph
// (dag.go:3076-3146) reports: an unknown dependency or a cycle.
func (c *Ctx) PutTopology(t *Topology) (uint32, error) {
	var col *C.uint16_t
	if len(t.ColIdx) > 0 {
		col = (*C.uint16_t)(unsafe.Pointer(&t.ColIdx[0]))
	}
	var par *C.bf_parallel_desc
	if len(t.Parallel) > 0 {
		par = (*C.bf_parallel_desc)(unsafe.Pointer(&t.Parallel[0]))
	}
	var slot C.uint32_t
	rc := C.bfgo_topology_put(c.p, C.uint32_t(len(t.StepFlags)), C.uint32_t(len(t.ColIdx)), u32p(t.RowPtr), col, u8p(t.StepFlags),
		par, C.uint32_t(len(t.Parallel)), u8p(t.AllowBits), C.uint32_t(8*len(t.AllowBits)), &slot)
	return uint32(slot), c.err(rc)
}

func (c *Ctx) DropTopology(slot uint32) error { return c.err(C.bf_topology_drop(c.p, C.uint32_t(slot))) }

// Batch is a pair of pinned C buffers (bf_alloc_pinned) laid out by bf_layout_init; the packer writes state
// records into State, the library writes result records into Result.  Both are C memory: no cgo pointer rule applies.
type Batch struct {
	Layout C.bf_layout
	N      uint32
	Cap    uint32
	State  unsafe.Pointer
	Result unsafe.Pointer
}

func (c *Ctx) NewBatch(stepsMax, childNibbles, fields, capRuns uint32) (*Batch, error) {
	b := &Batch{Cap: capRuns}
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

// StateRecord returns run r's state record as a byte slice over the pinned buffer.
func (b *Batch) StateRecord(r uint32) []byte {
	return unsafe.Slice((*byte)(unsafe.Add(b.State, uintptr(r)*uintptr(b.Layout.state_stride))), int(b.Layout.state_stride))
}

// ResultRecord returns run r's result record.
func (b *Batch) ResultRecord(r uint32) []byte {
	return unsafe.Slice((*byte)(unsafe.Add(b.Result, uintptr(r)*uintptr(b.Layout.result_stride))), int(b.Layout.result_stride))
}

// Counts mirrors bf_counts.
type Counts struct{ Ready, Skip, Expansion, Evals uint64 }

// Eval runs one frontier pass over the batch (H2D state -> kernels -> D2H results), synchronously and
// bounded; a non-nil error makes the reconciler requeue with backoff or fall back to the Go path.
func (c *Ctx) Eval(b *Batch, flags uint32) (Counts, error) {
	var k Counts
	rc := C.bfgo_eval(c.p, &b.Layout, C.uint32_t(b.N), C.uint32_t(flags), 0, b.State, b.Result, (*C.bf_counts)(unsafe.Pointer(&k)))
	return k, c.err(rc)
}

// Compact results: head[r] = low 15 summary bits | listed << 15 | event count << 16 (BF_HEAD_*), events = one uint16 per
// ready / skipped / failed step, step | kind << 10 (BF_EVT_*), run-major in batch order and step-ascending inside a run —
// the order of findReadySteps' lists.  Run r's events are the next head[r]>>16 entries.
const (
	HeadSummaryMask = 0x7FFF
	HeadListed      = 0x8000
	HeadCountShift  = 16
	EvalChangedOnly = 0x10 // BF_EVAL_CHANGED_ONLY: list only the runs whose result differs from the previous tick's
	// bf_eval_device only (a co-located GPU pipeline submitting passes over device buffers; the host-buffer calls ignore them):
	EvalCountsSet = 0x20 // BF_EVAL_COUNTS_SET: the pass overwrites the counts block, nobody zeroes it
	EvalPipelined = 0x40 // BF_EVAL_PIPELINED: the pass is independent of the preceding kernel of its stream and may start in its tail
)

// EvalCompact is Eval with the results as lists.
func (c *Ctx) EvalCompact(b *Batch, flags uint32, head []uint32, events []uint16) (nEvents uint64, nListed uint32, k Counts, err error) {
	var ev *C.uint16_t
	if len(events) > 0 {
		ev = (*C.uint16_t)(unsafe.Pointer(&events[0]))
	}
	var ne C.uint64_t
	var nl C.uint32_t
	rc := C.bfgo_eval_compact(c.p, &b.Layout, C.uint32_t(b.N), C.uint32_t(flags), 0, b.State, u32p(head), ev, C.uint64_t(len(events)),
		&ne, &nl, (*C.bf_counts)(unsafe.Pointer(&k)))
	return uint64(ne), uint32(nl), k, c.err(rc)
}

// SchedRun mirrors bf_sched_run (32 bytes, no pointers).
type SchedRun struct {
	StoryKey, QueueKey uint32
	Priority           int32
	QueuedElapsedS     uint32 // BF_SCHED_NONE when nothing is queued with a StartedAt
	RunPhase           uint32
	_                  [3]uint32
}

// SchedTables are the limits and what the batch does not hold (dag.go:1780-1961).
type SchedTables struct {
	StoryLimit           []int32  // Story.spec.policy.concurrency per story key
	StoryRunningBase     []uint32 // nil = zeros
	QueueLimit           []int32  // scheduling.queues[q].concurrency per queue key
	QueueAgingS          []int32  // scheduling.queues[q].priorityAgingSeconds
	QueueRunningBase     []uint32 // nil = zeros
	QueueMaxPriorityBase []int32  // highest effective priority among runs with demand outside the batch; nil = none
	GlobalLimit          int32    // scheduling.globalConcurrency
	GlobalBase           uint32
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
func (c *Ctx) Schedule(b *Batch, runs []SchedRun, t *SchedTables) (*SchedResult, error) {
	words := uint32(b.Layout.words)
	stride := (16 + 12*words + 15) &^ 15
	res := &SchedResult{
		Records:      make([]byte, int(b.N)*int(stride)),
		StoryRunning: make([]uint32, len(t.StoryLimit)),
		QueueRunning: make([]uint32, len(t.QueueLimit)),
		QueueMaxPrio: make([]int32, len(t.QueueLimit)),
	}
	var rp *C.bf_sched_run
	if len(runs) > 0 {
		rp = (*C.bf_sched_run)(unsafe.Pointer(&runs[0]))
	}
	var rec unsafe.Pointer
	if len(res.Records) > 0 {
		rec = unsafe.Pointer(&res.Records[0])
	}
	rc := C.bfgo_schedule(c.p, &b.Layout, C.uint32_t(b.N), rp, C.uint32_t(len(t.StoryLimit)), C.uint32_t(len(t.QueueLimit)),
		C.int32_t(t.GlobalLimit), C.uint32_t(t.GlobalBase), i32p(t.StoryLimit), u32p(t.StoryRunningBase), i32p(t.QueueLimit),
		i32p(t.QueueAgingS), u32p(t.QueueRunningBase), i32p(t.QueueMaxPriorityBase), rec, u32p(res.StoryRunning),
		u32p(res.QueueRunning), i32p(res.QueueMaxPrio), (*C.uint32_t)(unsafe.Pointer(&res.GlobalRunning)))
	return res, c.err(rc)
}

// RedriveClosure returns, for each (slot, step), the bit mask of the steps a redrive from that step resets
// (resolveRedriveFromStepSet, storyrun_controller.go:535-558).
func (c *Ctx) RedriveClosure(slots, steps []uint32, words uint32) ([]uint32, error) {
	masks := make([]uint32, len(slots)*int(words))
	if len(slots) == 0 {
		return masks, nil
	}
	return masks, c.err(C.bf_topology_closure(c.p, u32p(slots), u32p(steps), C.uint32_t(len(slots)), C.uint32_t(words), u32p(masks)))
}

// Delta mirrors bf_delta: one changed code of one state record.
type Delta struct {
	Run   uint32
	Index uint16
	Field uint8
	Code  uint8
}

// Resident is a device-resident batch (row f2): full records travel once, afterwards only deltas.
type Resident struct {
	c      *Ctx
	handle C.uint32_t
	layout C.bf_layout
}

func (c *Ctx) NewResident(layout C.bf_layout, capacity uint32) (*Resident, error) {
	r := &Resident{c: c, layout: layout}
	return r, c.err(C.bf_resident_create(c.p, &r.layout, C.uint32_t(capacity), &r.handle))
}

func (r *Resident) Close() error { return r.c.err(C.bf_resident_destroy(r.c.p, r.handle)) }

// Upload sends full state records for runs [first, first+n) (new StoryRuns, or a resync).
func (r *Resident) Upload(first, n uint32, records unsafe.Pointer) error {
	return r.c.err(C.bf_resident_upload(r.c.p, r.handle, C.uint32_t(first), C.uint32_t(n), records))
}

// Tick is the steady-state reconcile tick: the deltas syncStateFromStepRuns (dag.go:965-1009) produced go up, the
// pass runs over the resident state, and the ready / skipped steps come back as lists; with EvalChangedOnly only the
// runs whose result changed since the previous tick are listed (the batcher keeps every run's last Row).
func (r *Resident) Tick(deltas []Delta, n, flags uint32, head []uint32, events []uint16) (nEvents uint64, nListed uint32, k Counts, err error) {
	var dp *C.bf_delta
	if len(deltas) > 0 {
		dp = (*C.bf_delta)(unsafe.Pointer(&deltas[0]))
	}
	var ev *C.uint16_t
	if len(events) > 0 {
		ev = (*C.uint16_t)(unsafe.Pointer(&events[0]))
	}
	var ne C.uint64_t
	var nl C.uint32_t
	rc := C.bfgo_resident_tick_compact(r.c.p, r.handle, dp, C.uint32_t(len(deltas)), C.uint32_t(n), C.uint32_t(flags), 0, u32p(head), ev,
		C.uint64_t(len(events)), &ne, &nl, (*C.bf_counts)(unsafe.Pointer(&k)))
	return uint64(ne), uint32(nl), k, r.c.err(rc)
}

// Group is one operator process driving several GPUs (bf_group_*): contiguous blocks of runs per device, one NCCL
// all-gather of the per-shard counts per pass, limiter totals all-reduced across the shards.
type Group struct{ p *C.bf_group }

func NewGroup(devices []int32) (*Group, error) {
	var p *C.bf_group
	rc := C.bf_group_create(&p, i32p(devices), C.uint32_t(len(devices)), nil)
	if rc != C.BF_OK {
		return nil, &Error{Status: int(rc), Msg: "bf_group_create"}
	}
	return &Group{p: p}, nil
}

func (g *Group) Close()               { C.bf_group_destroy(g.p); g.p = nil }
func (g *Group) Size() uint32         { return uint32(C.bf_group_size(g.p)) }
func (g *Group) Shard(k uint32) *Ctx  { return &Ctx{p: C.bf_group_ctx(g.p, C.uint32_t(k))} }
func (g *Group) err(rc C.int) error {
	if rc == C.BF_OK {
		return nil
	}
	return &Error{Status: int(rc), Msg: C.GoString(C.bf_group_last_error(g.p))}
}

// ShardRange says which runs of an n-run batch live on shard k (so that the packer writes the shard's own slot ids).
func (g *Group) ShardRange(n, k uint32) (first, count uint32) {
	var f, cnt C.uint32_t
	C.bf_group_shard_range(g.p, C.uint32_t(n), C.uint32_t(k), &f, &cnt)
	return uint32(f), uint32(cnt)
}

// Eval runs the pass on every shard concurrently and returns the per-shard and the global counts.
func (g *Group) Eval(b *Batch, flags uint32) (perShard []Counts, global Counts, err error) {
	perShard = make([]Counts, g.Size())
	rc := C.bfgo_group_eval(g.p, &b.Layout, C.uint32_t(b.N), C.uint32_t(flags), 0, b.State, b.Result,
		(*C.bf_counts)(unsafe.Pointer(&perShard[0])), (*C.bf_counts)(unsafe.Pointer(&global)))
	return perShard, global, g.err(rc)
}
