package frontier

// FrontierProvider is the seam DAGReconciler gets (SOURCE ONLY, see frontier.go).
//
// internal/controller/runs/dag.go:1708 today:
//
//	readySteps, skippedSteps, skippedReasons, unskipped := r.findReadySteps(ctx, srun, story, steps,
//	    srun.Status.StepStates, completedSteps, runningSteps, dependencies, vars, depPolicy)
//
// becomes
//
//	if row, ok := r.Frontier.Lookup(srun.UID, srun.ResourceVersion); ok {
//	    readySteps, skippedSteps, skippedReasons = row.Steps(steps)   // bit i of ready/skip -> &steps[i]
//	} else {
//	    ... the Go path above (first sight of a run, packer miss, or bf_eval error)
//	}
//
// Ready bits are LSB-first in list order, so the concurrency limiters' readySteps[:slots] (dag.go:1796-1798)
// keep working on the slice returned by row.Steps.
type FrontierProvider interface {
	// Lookup returns the result row computed for this StoryRun at exactly this resourceVersion, if any.
	Lookup(uid string, resourceVersion string) (Row, bool)
}

// Row is one StoryRun's result record.
type Row struct {
	Summary           uint32
	Ready, Skip       []uint32 // bit masks, step i = bit i%32 of word i/32
	SkipDep, NeedCond []uint32
	PhaseOut          []uint32 // 4 bit planes
}
