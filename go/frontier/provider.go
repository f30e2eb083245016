package frontier

import (
	bubuv1alpha1 "github.com/bubustack/bobrapet/api/v1alpha1"
)

// FrontierProvider is the seam DAGReconciler gets (SOURCE ONLY, see frontier.go).
//
// internal/controller/runs/dag.go:1708 today:
//
//	readySteps, skippedSteps, skippedReasons, unskipped := r.findReadySteps(ctx, srun, story, steps,
//	    srun.Status.StepStates, completedSteps, runningSteps, dependencies, vars, depPolicy)
//
// becomes
//
//	if row, ok := r.Frontier.Lookup(string(srun.UID), srun.ResourceVersion); ok {
//	    readySteps, skippedSteps, skippedReasons = row.Steps(allStorySteps(story))
//	} else {
//	    ... the Go path above (first sight of a run, packer miss, or a bf_eval error)
//	}
//
// Ready bits are LSB-first in list order, so the concurrency limiters' readySteps[:slots] (dag.go:1796-1798)
// keep working on the slice returned by Row.Steps.
type FrontierProvider interface {
	// Lookup returns the result row computed for this StoryRun at exactly this resourceVersion, if any.
	Lookup(uid string, resourceVersion string) (Row, bool)
}

// Row is one StoryRun's result of the last tick that listed it, decoded from the compact lists.
type Row struct {
	Summary uint32   // low 15 bits of BF_SUM_*: group evaluated, mainDone / mainFailed flags, "some phase changed"
	Events  []uint16 // this run's slice of the tick's event list: step | kind << 10
	// FailedDep names, per skipped step, the dependency that failed — "Skipped due to failed dependency: <d>"
	// (dag.go:2735-2739): the lowest index among the step's needs whose phase is terminal and not Succeeded/Skipped
	// (the packer fills it from the host copy of the phases; the kernel reports only WHICH steps were skipped for it).
	FailedDep map[uint16]string
}

const (
	evtReady   = 0x1 // BF_EVT_READY
	evtSkip    = 0x2 // BF_EVT_SKIP
	evtSkipDep = 0x10
)

// Rows walks one tick's lists: run r's events are the next head[r]>>16 entries; only listed runs are handed to `set`
// (with BF_EVAL_CHANGED_ONLY the cache keeps the previous Row of every other run).
func Rows(head []uint32, events []uint16, set func(run uint32, row Row)) {
	pos := 0
	for r, h := range head {
		n := int(h >> HeadCountShift)
		if h&HeadListed != 0 {
			set(uint32(r), Row{Summary: h & HeadSummaryMask, Events: events[pos : pos+n]})
		}
		pos += n
	}
}

// Steps turns the row back into the three results of findReadySteps (dag.go:2631-2641) over `steps` =
// allStorySteps(story) — index i of the packed topology is &steps[i].
func (r Row) Steps(steps []bubuv1alpha1.Step) (ready, skipped []*bubuv1alpha1.Step, reasons map[string]string) {
	reasons = map[string]string{}
	for _, e := range r.Events {
		step, kind := e&0x3FF, e>>10
		if int(step) >= len(steps) {
			continue
		}
		st := &steps[step]
		switch {
		case kind&evtReady != 0:
			ready = append(ready, st)
		case kind&evtSkip != 0:
			skipped = append(skipped, st)
			if kind&evtSkipDep != 0 || r.FailedDep[step] != "" {
				reasons[st.Name] = "Skipped due to failed dependency: " + r.FailedDep[step] // dag.go:2736
			} else {
				reasons[st.Name] = "Skipped due to 'if' condition" // dag.go:2831
			}
		}
	}
	return ready, skipped, reasons
}
