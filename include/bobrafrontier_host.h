/*
 * bobrafrontier_host.h — host-side mirror of the reference's objects for the frontier path.
 *
 * What the Go batcher would otherwise re-implement: turning Story / StoryRun objects (names, `needs`, template
 * strings, StepState phases and messages, GateStatus, child StepRuns) into the packed records of
 * bobrafrontier.h, and result records back into step names.  Reference functions mirrored (paths relative to
 * /root/reference/internal/controller/runs):
 *
 *   buildDependencyGraphs / findAndAddDeps / addDependency   dag.go:3024-3073, 3223-3268  (bfh_story_finalize)
 *   sanitizeStepIdentifier (alias map)                       step_executor.go:1652-1670
 *   validateRuntimeDependencyGraph (unknown deps)            dag.go:3076-3098            (cycles: bf_topology_put)
 *   allStorySteps (index space main ++ comp ++ finally)      dag.go:3270-3280
 *   isConcurrencyQueued (message prefixes -> code 14)        dag.go:2035-2051            (bfh_run_set_phase)
 *   checkSyncGates decision order                            dag.go:1489-1533            (bfh_run_set_gate)
 *   syncStateFromStepRuns-style incremental updates          dag.go:965-1009             (bfh_run_set_* are O(1))
 *
 * State records live in pinned host memory owned by the batch and are updated IN PLACE (a phase change flips at
 * most four bits), so a tick's host work is O(changes) and bf_eval's upload is one contiguous copy.
 *
 * Same ABI rules as bobrafrontier.h: flat C, negative bf_status on error, nothing retained but what is documented.
 */
#ifndef BOBRAFRONTIER_HOST_H_
#define BOBRAFRONTIER_HOST_H_

#include "bobrafrontier.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ Story (definition side) */
typedef struct bfh_story bfh_story;

bfh_story* bfh_story_new(void);
void bfh_story_free(bfh_story* s);
const char* bfh_story_error(const bfh_story* s);

/* Append a step in list order within its group.  group: BF_GROUP_MAIN / _COMPENSATION / _FINALLY;
 * type: BF_STEP_* (BF_STEP_ENGRAM = step.Ref set).  if_expr / with_raw may be NULL.  Returns the step's
 * insertion handle (>= 0) or a negative status. */
int bfh_story_add_step(bfh_story* s, const char* name, int group, int type, int allow_failure, int on_timeout_skip,
                       const char* if_expr, const char* with_raw);
int bfh_step_add_need(bfh_story* s, int step, const char* dep_name);
int bfh_step_add_branch(bfh_story* s, int step, const char* branch_name, int allow_failure); /* parallel: with.steps */
/* continue_on_step_failure: -1 unset, 0, 1 (Policy.Retries.ContinueOnStepFailure, dag.go:3504-3511) */
int bfh_story_set_policy(bfh_story* s, int continue_on_step_failure, int realtime);

/* Resolve names, extract template references, build the CSR over allStorySteps.  BF_ETOPO on unknown deps. */
int bfh_story_finalize(bfh_story* s);
int bfh_story_dims(const bfh_story* s, uint32_t* n_steps, uint32_t* n_edges, uint32_t* n_parallel);
/* Copy out the packed form (row_ptr[S+1], col_idx[E], step_flags[S]); any pointer may be NULL. */
int bfh_story_csr(const bfh_story* s, uint32_t* row_ptr, uint16_t* col_idx, uint8_t* step_flags);
int bfh_story_upload(bfh_story* s, bf_ctx* ctx, uint32_t* slot_out); /* bf_topology_put of the packed form */
int bfh_story_step_index(const bfh_story* s, const char* name);      /* index in allStorySteps order, or -1 */
const char* bfh_story_step_name(const bfh_story* s, uint32_t index);
uint32_t bfh_story_run_flags(const bfh_story* s);                     /* BF_RF_FAIL_FAST / BF_RF_REALTIME from the policy */

/* One pass of the template-reference scanner (exposed for tests): writes up to cap names, '\n'-separated. */
int bfh_scan_step_refs(const char* expression, char* out, size_t cap);

/* ------------------------------------------------------------------ Batch of live StoryRuns */
typedef struct bfh_batch bfh_batch;

/* ctx may be NULL (CPU-side packing only: bfh_batch_eval then returns BF_ENODEV). */
bfh_batch* bfh_batch_new(bf_ctx* ctx, uint32_t steps_max, uint32_t child_nibbles, uint32_t fields, uint32_t capacity);
void bfh_batch_free(bfh_batch* b);
const char* bfh_batch_error(const bfh_batch* b);
const bf_layout* bfh_batch_layout(const bfh_batch* b);
uint32_t bfh_batch_size(const bfh_batch* b);
const void* bfh_batch_state(const bfh_batch* b);   /* [size * state_stride]  */
const void* bfh_batch_result(const bfh_batch* b);  /* [size * result_stride] */

/* Add a StoryRun of `story` (finalized; `slot` from bfh_story_upload or any caller-chosen slot id).  Returns the
 * run index.  The run starts with no StepStates, run flags from the Story policy. */
int bfh_batch_add_run(bfh_batch* b, const bfh_story* story, uint32_t slot);
int bfh_batch_remove_last_run(bfh_batch* b);

/* StepState update.  phase: "", "Pending", ... ; message only matters for Pending (queued prefixes -> code 14). */
int bfh_run_set_phase(bfh_batch* b, uint32_t run, uint32_t step, const char* phase, const char* message);
int bfh_run_set_phase_code(bfh_batch* b, uint32_t run, uint32_t step, int code);
int bfh_run_set_cond(bfh_batch* b, uint32_t run, uint32_t step, int cond_code);         /* BF_COND_* */
int bfh_run_set_decision(bfh_batch* b, uint32_t run, uint32_t step, int decision_code); /* BF_DEC_*  */
/* Gate decision exactly as checkSyncGates orders it: Approved > Rejected > timed-out > pending. */
int bfh_run_set_gate(bfh_batch* b, uint32_t run, uint32_t step, const char* gate_state, int timed_out);
int bfh_run_set_run_flags(bfh_batch* b, uint32_t run, int topology_terminated, int host_group /* -1 none, else BF_GROUP_* */);
int bfh_run_register_children(bfh_batch* b, uint32_t run, uint32_t parallel_index, int registered);
int bfh_run_set_child_phase(bfh_batch* b, uint32_t run, uint32_t parallel_index, uint32_t branch, const char* phase);

/* One frontier pass over the batch through bf_eval (eval_flags: BF_EVAL_*). */
int bfh_batch_eval(bfh_batch* b, uint32_t eval_flags, bf_counts* counts);

/* Results of the last pass.  Step lists come back in list order (ascending index), the order the
 * concurrency limiters' readySteps[:slots] relies on (dag.go:1796-1798). */
uint32_t bfh_run_summary(const bfh_batch* b, uint32_t run);
int bfh_run_ready(const bfh_batch* b, uint32_t run, uint32_t* steps_out, uint32_t cap);
int bfh_run_skipped(const bfh_batch* b, uint32_t run, uint32_t* steps_out, uint32_t cap);
int bfh_run_failed(const bfh_batch* b, uint32_t run, uint32_t* steps_out, uint32_t cap);
int bfh_run_needs_cond(const bfh_batch* b, uint32_t run, uint32_t* steps_out, uint32_t cap);
int bfh_run_phase_out(const bfh_batch* b, uint32_t run, uint32_t step); /* phase code after the pass, or -1 */
/* "Skipped due to failed dependency: <name>" names the first dep in CSR order that is failed. */
int bfh_run_skip_reason(const bfh_batch* b, uint32_t run, uint32_t step, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BOBRAFRONTIER_HOST_H_ */
