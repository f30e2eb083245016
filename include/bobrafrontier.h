/*
 * bobrafrontier.h — C ABI of the B200-native StoryRun DAG ready-frontier engine.
 *
 * This is the drop-in boundary for ONE hot path of bubustack/bobrapet: the
 * per-reconcile "which steps are ready now?" computation
 *
 *     DAGReconciler.findReadySteps      internal/controller/runs/dag.go:2631-2848
 *     buildStateMaps                    internal/controller/runs/dag.go:3358-3391
 *     clearConcurrencyQueuedSteps       internal/controller/runs/dag.go:2020-2051
 *     checkSyncGates/Sleep/Wait         internal/controller/runs/dag.go:1455-1547, 1217-1288, 1291-1452
 *     checkSyncParallelSteps            internal/controller/runs/dag.go:1112-1200
 *     group selection, fail-fast,       internal/controller/runs/dag.go:422-511, 3289-3342
 *       compensation skipping
 *     launch effects of Execute         internal/controller/runs/step_executor.go:132-185,
 *                                       internal/controller/runs/dag.go:1735-1775
 *
 * evaluated for a whole BATCH of live StoryRuns in one pass on the GPU.  The
 * reference has no FFI for this path (it is Go-internal); these entry points
 * are what a cgo shim (go/frontier/frontier.go) binds.
 *
 * Rules of the ABI: flat PODs, caller-owned buffers, no callbacks, no C++
 * types, never aborts, never throws; every function returns 0 or a negative
 * bf_status.  The library retains nothing the caller passed except the
 * topologies (copied to HBM by bf_topology_put*).
 *
 * Everything here is integer/bit work; results are bit-exact against the
 * oracle (oracle/), see DESIGN.md.
 */
#ifndef BOBRAFRONTIER_H_
#define BOBRAFRONTIER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_ABI_VERSION 3u

/* ------------------------------------------------------------------ status */
typedef enum bf_status {
  BF_OK = 0,
  BF_EINVAL = -1, /* bad argument / malformed batch                              */
  BF_ENOMEM = -2, /* host or device allocation failed                            */
  BF_ECUDA = -3,  /* CUDA runtime error (bf_last_error has the text)             */
  BF_ENCCL = -4,  /* collective error (multi-GPU count exchange)                 */
  BF_ETOPO = -5,  /* topology rejected: unknown dep, cycle, too large, bad slot  */
  BF_ENODEV = -6  /* no usable sm_100 device / extension built without CUDA      */
} bf_status;

/* ---------------------------------------------------------- phase encoding */
/* 4-bit code per (run, step), stored bit-sliced (see "record layout").  Order = declaration order of enums.Phase
 * (pkg/enums/enums.go:44-97); 0 = "no StepState entry".  Code 14 folds the
 * "queued" bit into the nibble: phase Pending whose message starts with one of
 * the four "Queued due to ..." prefixes (dag.go:103-108, isConcurrencyQueued
 * dag.go:2035-2051).  It is Pending for every purpose except that
 * clearConcurrencyQueuedSteps removes it from the running set.               */
enum {
  BF_PHASE_NONE = 0,
  BF_PHASE_PENDING = 1,
  BF_PHASE_RUNNING = 2,
  BF_PHASE_SUCCEEDED = 3,
  BF_PHASE_FAILED = 4,
  BF_PHASE_FINISHED = 5,
  BF_PHASE_CANCELED = 6,
  BF_PHASE_COMPENSATED = 7,
  BF_PHASE_PAUSED = 8,
  BF_PHASE_BLOCKED = 9,
  BF_PHASE_SCHEDULING = 10,
  BF_PHASE_TIMEOUT = 11,
  BF_PHASE_ABORTED = 12,
  BF_PHASE_SKIPPED = 13,
  BF_PHASE_PENDING_QUEUED = 14,
  BF_PHASE_RESERVED = 15 /* invalid on input; rejected by BF_EVAL_VALIDATE */
};

/* 16-entry boolean tables over the phase code (bit v = value for code v).   */
#define BF_LUT_TERMINAL 0x38F8u   /* Phase.IsTerminal, enums.go:101-115          */
#define BF_LUT_COMPLETED0 0x2008u /* Succeeded|Skipped, dag.go:3377              */
#define BF_LUT_RUNNING 0x4106u    /* Running|Pending|Paused (+queued), :3379     */
#define BF_LUT_RUNNING_Q 0x0106u  /* ... minus concurrency-queued, :2020-2033    */
#define BF_LUT_RT_SAT 0x410Eu     /* dependencySatisfiedForRealtime, :3457-3473  */

/* ------------------------------------------------------------- step flags  */
/* One byte per step, static per Story generation (bf_topology_put).         */
enum {
  BF_STEP_ENGRAM = 0, /* step.Ref != nil                         */
  BF_STEP_CONDITION = 1,
  BF_STEP_PARALLEL = 2,
  BF_STEP_SLEEP = 3,
  BF_STEP_STOP = 4,
  BF_STEP_WAIT = 5,
  BF_STEP_EXECUTE_STORY = 6,
  BF_STEP_GATE = 7
};
#define BF_SF_TYPE_MASK 0x07u
#define BF_SF_ALLOW_FAILURE 0x08u   /* step.AllowFailure, story_types.go:193      */
#define BF_SF_ON_TIMEOUT_SKIP 0x10u /* with.onTimeout == "skip", dag.go:1655-1668 */
#define BF_SF_HAS_IF 0x20u          /* step.If != nil && != "", dag.go:2741       */
#define BF_SF_GROUP_SHIFT 6         /* 0 main (spec.steps), 1 compensations, 2 finally */
#define BF_SF_GROUP_MASK 0xC0u
enum { BF_GROUP_MAIN = 0, BF_GROUP_COMPENSATION = 1, BF_GROUP_FINALLY = 2, BF_GROUP_DONE = 3 };

/* -------------------------------------------------------------- run flags  */
#define BF_RF_FAIL_FAST 0x01u           /* shouldFailFast, dag.go:3504-3511             */
#define BF_RF_REALTIME 0x02u            /* story.Spec.Pattern.IsRealtime()              */
#define BF_RF_TOPOLOGY_TERMINATED 0x04u /* Degraded/TopologyTerminated, dag.go:436-464  */
#define BF_RF_HOST_GROUP 0x08u          /* tier K1: group given by host, skip stage I   */
#define BF_RF_HOST_GROUP_SHIFT 4        /* bits 4-5: BF_GROUP_* when BF_RF_HOST_GROUP   */

/* ----------------------------------------------------- cond / decision codes */
/* 2 bits per (run, step), stored as 2 bit planes.                            */
enum {
  BF_COND_PASS = 0, /* no `if`, realtime, or evaluated true and `with` refs fresh (dag.go:2741,2837,2844) */
  BF_COND_SKIP = 1, /* evaluated false and refs not stale (dag.go:2820-2832)                            */
  BF_COND_HOLD = 2, /* blocked / eval error / offloaded wait / stale refs (dag.go:2801-2842)            */
  BF_COND_FAIL = 3  /* template-safety violation or offloaded under fail policy (dag.go:2744,2810)      */
};
enum {
  BF_DEC_PENDING = 0,   /* keep waiting -> Paused                                      */
  BF_DEC_SUCCEED = 1,   /* gate approved / sleep elapsed / wait satisfied -> Succeeded  */
  BF_DEC_FAIL = 2,      /* gate rejected / invalid config / fail policy  -> Failed      */
  BF_DEC_TIMED_OUT = 3  /* timeout elapsed while pending -> Skipped|Timeout (dag.go:1655) */
};

/* ---------------------------------------------------------------- limits   */
#define BF_MAX_STEPS 1024u   /* steps per topology (CRD caps at 100+50+50, story_types.go:129-140) */
#define BF_MAX_EDGES 65535u  /* u16 row_ptr on device                                             */
#define BF_MAX_PARALLEL 64u  /* parallel steps per topology (children_registered is a u64)        */
#define BF_OFF_NONE 0xFFFFFFFFu

/* ------------------------------------------------------------- topologies  */
typedef struct bf_parallel_desc {
  uint16_t step;        /* index of the `parallel` step                                       */
  uint16_t branches;    /* B = len(with.steps), step_executor.go:745-806                      */
  uint32_t allow_first; /* index of branch 0's bit in branch_allow_bits (bit i -> branch i)   */
} bf_parallel_desc;

typedef struct bf_topology {
  uint32_t n_steps;                /* S: main ++ compensations ++ finally (allStorySteps, dag.go:3270) */
  uint32_t n_edges;                /* E                                                               */
  const uint32_t* row_ptr;         /* [S+1], row_ptr[0]=0, row_ptr[S]=E; deps of step i = col_idx[row_ptr[i]..row_ptr[i+1]) */
  const uint16_t* col_idx;         /* [E] dependency step indices (buildDependencyGraphs, dag.go:3024-3073) */
  const uint8_t* step_flags;       /* [S] BF_SF_*                                                     */
  const bf_parallel_desc* parallel; /* [n_parallel], ascending by step                                */
  uint32_t n_parallel;
  const uint8_t* branch_allow_bits; /* branch allowFailure bits (dag.go:1145-1158); may be NULL       */
  uint32_t n_branch_allow_bits;
} bf_topology;

/* ------------------------------------------------------------ record layout */
/* Dynamic state travels as ONE record per run (so the host->device copy is a
 * single contiguous transfer and the kernel stages a run with one bulk copy).
 * Per-step codes are BIT-SLICED: a k-bit code is stored as k bit planes of
 * `words` u32 each, step i = bit (i%32) of word i/32 of every plane, plane b
 * holding bit b of the code.  Same size as nibble packing (4 bits/step for the
 * phase), but 32 steps are classified by a handful of bitwise ops and no
 * transposition is ever needed on either side.
 *
 *   offset 0     bf_run_header (16 B)
 *   off_phase    phase code, 4 planes: plane b at off_phase + b*words*4     (words*16 B)
 *   off_cond     cond code, 2 planes                                        (words*8 B)  optional
 *   off_decision decision code, 2 planes                                    (words*8 B)  optional
 *   off_child    child StepRun phase NIBBLES of parallel steps (branch order), desc p
 *                starts at nibble child_first[p] (assigned by bf_topology_put),
 *                child c at byte c/2, low nibble first                                   optional
 *
 * Results likewise, one record per run:
 *
 *   offset 0   bf_result_header (16 B)
 *   off_ready / off_skip / off_fail / off_needs_cond / off_skip_dep  bit masks,
 *              step i = bit (i%32) of u32 word i/32  (words*4 B each)
 *   off_phase_out  phase code after this pass's status mutations, 4 planes  (words*16 B)
 *
 * All strides are multiples of 16 bytes.                                      */
typedef struct bf_run_header {
  uint32_t topo_slot;           /* from bf_topology_put                                          */
  uint8_t run_flags;            /* BF_RF_*                                                       */
  uint8_t reserved[3];
  uint64_t children_registered; /* bit p: PrimitiveChildren[parallel p] non-empty, dag.go:1140   */
} bf_run_header;

typedef struct bf_result_header {
  uint32_t summary;      /* BF_SUM_*                                                             */
  uint32_t n_ready;      /* popcount(ready)                                                      */
  uint32_t n_skip;       /* popcount(skip)                                                       */
  uint32_t n_expansion;  /* child StepRuns to create for ready `parallel` steps                   */
} bf_result_header;

#define BF_SUM_GROUP_MASK 0x3u       /* BF_GROUP_* evaluated this pass (3 = finalize on host, dag.go:490-495) */
#define BF_SUM_MAIN_DONE 0x4u        /* dag.go:431                                                        */
#define BF_SUM_MAIN_FAILED 0x8u      /* len(mainFailed) > 0                                               */
#define BF_SUM_COMP_DONE 0x10u
#define BF_SUM_FINAL_DONE 0x20u
#define BF_SUM_COMP_FAILED 0x40u
#define BF_SUM_FINAL_FAILED 0x80u
#define BF_SUM_PHASE_CHANGED 0x100u  /* some phase nibble was rewritten (needsPersist, dag.go:409-470)    */
#define BF_SUM_ITER_SHIFT 16         /* bits 16-27: iterations run (fixpoint mode), dag.go:393            */

/* field selection for bf_layout_init */
#define BF_F_COND 0x1u
#define BF_F_DECISION 0x2u
#define BF_F_CHILD 0x4u
#define BF_F_OUT_FAIL 0x10u
#define BF_F_OUT_NEEDS_COND 0x20u
#define BF_F_OUT_SKIP_DEP 0x40u
#define BF_F_OUT_PHASE 0x80u

typedef struct bf_layout {
  uint32_t steps_max; /* max S over the topologies of the batch                      */
  uint32_t words;     /* ceil(steps_max/32)                                          */
  uint32_t fields;    /* BF_F_*                                                      */
  uint32_t child_nibbles;
  uint32_t state_stride;
  uint32_t off_phase, off_cond, off_decision, off_child;
  uint32_t result_stride;
  uint32_t off_ready, off_skip, off_fail, off_needs_cond, off_skip_dep, off_phase_out;
} bf_layout;

/* ------------------------------------------------------------------ batch  */
#define BF_EVAL_VALIDATE 0x1u   /* host API only: check slots / reserved nibbles before launch   */
#define BF_EVAL_FIXPOINT 0x2u   /* iterate runDagIterations on device (dag.go:393-540), applying
                                   launch effects of instantly-completing primitives             */
#define BF_EVAL_EXPANSION 0x4u  /* emit (run, step, branch) tuples for ready parallel steps      */
#define BF_EVAL_NO_COUNTS 0x8u  /* skip the global counters                                      */
/* bf_eval_device only (0x10 is BF_EVAL_CHANGED_ONLY below):                                                          */
#define BF_EVAL_COUNTS_SET 0x20u /* counts are OVERWRITTEN with this pass's totals (no zeroing by the caller, no memset
                                   launch between passes: the last CTA out publishes the totals)                      */
#define BF_EVAL_PIPELINED 0x40u  /* caller's promise: this pass reads nothing the preceding kernel of the stream writes
                                   and writes nothing it reads or writes (other batch, other result / counts buffers).
                                   The pass is then launched as a programmatic dependent of that kernel: its CTAs start
                                   on SMs the previous pass has already left, so its start-up (slot ids -> slot entries
                                   -> first bulk copies, ~6 us) overlaps the previous pass's tail.  Completion order is
                                   kept (the pass completes after the preceding kernel).  Needs BF_EVAL_COUNTS_SET or
                                   BF_EVAL_NO_COUNTS; ignored where it does not apply                               */

typedef struct bf_expansion {
  uint32_t run;    /* index in the batch                      */
  uint16_t step;   /* the parallel step                       */
  uint16_t branch; /* branch index in with.steps              */
} bf_expansion;

typedef struct bf_counts {
  uint64_t ready;     /* sum n_ready over the batch                                   */
  uint64_t skip;      /* sum n_skip                                                   */
  uint64_t expansion; /* sum n_expansion                                              */
  uint64_t evals;     /* (run, step) pairs visited = sum of S over runs (dag.go:2647) */
} bf_counts;

typedef struct bf_batch {
  uint32_t struct_size; /* sizeof(bf_batch)                                            */
  uint32_t n_runs;
  uint32_t flags;       /* BF_EVAL_*                                                   */
  uint32_t max_iterations; /* fixpoint cap; 0 = S+1 (dag.go:393)                        */
  bf_layout layout;
  const void* state;    /* [n_runs * state_stride]                                     */
  void* result;         /* [n_runs * result_stride]                                    */
  bf_expansion* expansion; /* [expansion_cap], run-major/step/branch order; may be NULL */
  uint64_t expansion_cap;
  bf_counts* counts;    /* out; may be NULL                                            */
} bf_batch;

/* ------------------------------------------------------------------- ctx   */
typedef struct bf_ctx bf_ctx;

typedef struct bf_config {
  uint32_t struct_size;
  int32_t device;          /* CUDA ordinal; one ctx drives one GPU (one process per GPU) */
  uint64_t arena_bytes;    /* initial topology arena in HBM (grows on demand); 0 = default */
  uint32_t max_topologies; /* initial slot table size; 0 = default                       */
  uint32_t flags;          /* BF_CFG_*                                                   */
} bf_config;
/* The frontier kernels are persistent (one CTA per SM holding nearly all of its shared memory), so a kernel of ANOTHER
 * stream — the NCCL all-gather of the counts — finds no SM until a pass ends and queues behind it.  Leaving k SMs out of
 * the grid (k = flags & 0xFF) lets the collective run beside the next pass; it costs k / 148 of the pass.                */
#define BF_CFG_RESERVE_SMS(k) ((uint32_t)(k) & 0xFFu)

uint32_t bf_abi_version(void);
const char* bf_strerror(int status);
/* Last error text of this ctx (valid until the next call on the ctx). */
const char* bf_last_error(const bf_ctx* ctx);

int bf_create(bf_ctx** out, const bf_config* cfg);
void bf_destroy(bf_ctx* ctx);

/* Upload one topology; returns its slot in *slot_out.  Validates what
 * validateRuntimeDependencyGraph (dag.go:3076-3146) validates: every col_idx
 * < S (no unknown dep) and acyclicity (Kahn), else BF_ETOPO.                  */
int bf_topology_put(bf_ctx* ctx, const bf_topology* topo, uint32_t* slot_out);
/* Bulk form: `count` topologies, one bf_topology each; slots_out[count].      */
int bf_topology_put_many(bf_ctx* ctx, const bf_topology* topos, uint32_t count, uint32_t* slots_out);
/* Device-validated bulk upload: unknown deps are still checked on the host, acyclicity is
 * checked by a kernel over the uploaded records (one warp per topology, level-synchronous peeling) instead of
 * the host's Kahn pass.  status_out[i]: bit 0 = cycle detected (the topology is dropped, slots_out[i] =
 * 0xFFFFFFFF), bits 8.. = number of dependency levels.  Returns BF_OK even when some topologies are cyclic. */
int bf_topology_put_many_checked_on_device(bf_ctx* ctx, const bf_topology* topos, uint32_t count, uint32_t* slots_out,
                                           uint32_t* status_out);
/* Re-validate already uploaded topologies on the device (status words as above; 0xFFFFFFFF = unknown slot). */
int bf_topology_check(bf_ctx* ctx, const uint32_t* slots, uint32_t count, uint32_t* status_out);

/* Redrive closure (row f3): for each query (topology slot, step index) the set of steps a redrive from that step
 * resets — the step plus everything downstream of it inside its own group (resolveRedriveFromStepSet,
 * internal/controller/runs/storyrun_controller.go:535-558; findStepGroup :560-577).  masks_out[q][words], same bit
 * order as the ready masks.  BF_EINVAL for a step index >= S ("step %q not found", :541).  Host buffers, synchronous. */
int bf_topology_closure(bf_ctx* ctx, const uint32_t* slots, const uint32_t* steps, uint32_t count, uint32_t words,
                        uint32_t* masks_out);

int bf_topology_drop(bf_ctx* ctx, uint32_t slot);
/* Nibble offset of parallel desc p's children inside the child area.          */
int bf_topology_child_first(const bf_ctx* ctx, uint32_t slot, uint32_t* child_first_out, uint32_t cap);

int bf_layout_init(bf_layout* out, uint32_t steps_max, uint32_t child_nibbles, uint32_t fields);

/* Synchronous pass over HOST buffers: H2D(state) -> kernels -> D2H(result),
 * returns when result/expansion/counts are filled.  Thread-safe per ctx.      */
int bf_eval(bf_ctx* ctx, const bf_batch* batch);

/* Asynchronous pass over DEVICE buffers on `stream` (a cudaStream_t): state,
 * result, expansion and counts are device pointers; nothing is synchronised.
 * counts must be zeroed by the caller unless BF_EVAL_NO_COUNTS or
 * BF_EVAL_COUNTS_SET (without them the library only adds).                    */
int bf_eval_device(bf_ctx* ctx, const bf_batch* batch, void* stream);

/* ------------------------------------------------------------------ limiters
 * The consumer of the ready sets: findAndLaunchReadySteps applies enforceStoryConcurrency (dag.go:1780-1799)
 * and enforceSchedulingLimits (:1801-1861, with enforcePriorityOrdering :1910-1946 and effectivePriority
 * :1948-1961) to each run's ready LIST and keeps a prefix (readySteps[:slots]); the rest is marked
 * "Queued due to ...".  Ready masks are LSB-first = list order, so the prefix is "the first k set bits".
 *
 * Batch contract (one consistent snapshot per tick): first the counts the reference obtains from cluster LISTs
 * are REDUCED over the batch on the device —
 *   running StepRuns of a run = engram steps in phase Running + Running children of registered parallel steps
 *                               (StepState mirrors StepRun.Status.Phase after syncStateFromStepRuns, :965-1009),
 *   per story key / per queue key / global totals = host-supplied base (StepRuns the batch does not hold) + sum,
 *   per queue the highest effective priority among non-terminal runs with demand (storyRunHasDemand :1981-1999:
 *   run phase Running/Pending, or a step Running or queued) —
 * then every run is limited independently against them (StepRuns created in the same tick have no phase yet
 * and do not count, exactly as in the reference).  Host work that stays on the host: label -> key mapping,
 * time (queued_elapsed_s = int32((now - storyRunQueuedSince).Seconds()), BF_SCHED_NONE when nothing is queued
 * with a StartedAt), formatting "(%d running, limit %d)" from the totals returned here.
 * A run's own label priority is the priority its Story resolves to (ensureSchedulingLabels, scheduling.go:48-70).
 */
#define BF_SCHED_NONE 0xFFFFFFFFu
enum {                       /* why the steps in queued_sched were queued                               */
  BF_QUEUED_NONE = 0,
  BF_QUEUED_PRIORITY = 1,    /* "Queued due to higher-priority work"      dag.go:1807-1811               */
  BF_QUEUED_GLOBAL = 2,      /* "Queued due to global concurrency limit"  dag.go:1849-1850               */
  BF_QUEUED_QUEUE = 3,       /* "Queued due to queue concurrency limit"   dag.go:1851-1852               */
  BF_QUEUED_OTHER = 4        /* "Queued due to scheduling limits"         dag.go:1853-1855               */
};

typedef struct bf_sched_run { /* 32 B per run                                                           */
  uint32_t story_key;         /* index into the story tables (namespace + Story name)                    */
  uint32_t queue_key;         /* index into the queue tables (queueLabelValue of the scheduling decision)*/
  int32_t priority;           /* schedulingDecision.Priority == the run's priority label                 */
  uint32_t queued_elapsed_s;  /* seconds queued, BF_SCHED_NONE if no queued step has a StartedAt         */
  uint32_t run_phase;         /* BF_PHASE_* of StoryRun.Status.Phase                                     */
  uint32_t reserved[3];
} bf_sched_run;

typedef struct bf_sched_tables {
  uint32_t struct_size;
  uint32_t n_stories, n_queues;
  int32_t global_limit;                /* scheduling.globalConcurrency, <= 0: none                       */
  uint32_t global_running_base;
  const int32_t* story_limit;          /* [n_stories] Story.spec.policy.concurrency, <= 0: none          */
  const uint32_t* story_running_base;  /* [n_stories] or NULL (zeros)                                     */
  const int32_t* queue_limit;          /* [n_queues]  queues[q].concurrency, <= 0: none                  */
  const int32_t* queue_aging_s;        /* [n_queues]  queues[q].priorityAgingSeconds, <= 0: no aging      */
  const uint32_t* queue_running_base;  /* [n_queues] or NULL (zeros)                                      */
  const int32_t* queue_max_priority_base; /* [n_queues] highest effective priority among non-terminal runs with demand that the
                                          batch does NOT hold (other shards, runs the batcher left out), INT32_MIN where none;
                                          NULL = none anywhere.  enforcePriorityOrdering (dag.go:1917-1944) compares against
                                          EVERY non-terminal StoryRun of the queue, so a sharded / partial batch must be given
                                          the rest here (bf_group_schedule does the exchange between shards itself)          */
} bf_sched_tables;

typedef struct bf_sched_header { /* first 16 B of a schedule record                                     */
  uint32_t n_launch;             /* popcount(launch): steps to hand to StepExecutor.Execute              */
  uint32_t n_queued_story;       /* queued by the Story's own concurrency limit                          */
  uint32_t n_queued_sched;       /* queued by priority ordering / queue / global limits                  */
  uint32_t sched_reason;         /* BF_QUEUED_*                                                          */
} bf_sched_header;

/* schedule record of a run: bf_sched_header, then launch[words], queued_story[words], queued_sched[words]
 * (u32 bit masks, same bit order as ready); stride = BF_SCHED_STRIDE(words), a multiple of 16.           */
#define BF_SCHED_STRIDE(words) ((16u + 12u * (uint32_t)(words) + 15u) & ~15u)

typedef struct bf_sched_out {
  uint32_t struct_size;
  uint32_t reserved;
  void* records;                 /* [n_runs][BF_SCHED_STRIDE(layout.words)]                               */
  uint32_t* story_running;       /* [n_stories] totals used (base + batch), or NULL                       */
  uint32_t* queue_running;       /* [n_queues] or NULL                                                    */
  int32_t* queue_max_priority;   /* [n_queues] INT32_MIN when no run of the queue has demand, or NULL     */
  uint32_t* global_running;      /* [1] or NULL                                                           */
} bf_sched_out;

/* Host buffers.  Limits the ready sets of the batch evaluated by the IMMEDIATELY PRECEDING bf_eval on this
 * ctx (same n_runs and layout; its state and result records are still on the device).  Synchronous.      */
int bf_schedule(bf_ctx* ctx, const bf_batch* batch, const bf_sched_run* runs, const bf_sched_tables* tables,
                bf_sched_out* out);
/* Device buffers: batch->state / batch->result as for bf_eval_device, runs / table arrays / out arrays are
 * device pointers (totals arrays are required here: they are the reduction scratch).  Asynchronous on `stream`. */
int bf_schedule_device(bf_ctx* ctx, const bf_batch* batch, const bf_sched_run* runs, const bf_sched_tables* tables,
                       bf_sched_out* out, void* stream);


/* ------------------------------------------------------------------ resident batches
 * Incremental state upload: the state records of a batch stay on the device and the host sends DELTAS — what
 * syncStateFromStepRuns (dag.go:965-1009) and the StepRun watch events (storyrun_controller.go:2116-2129) amount
 * to per tick: a handful of (run, step) phase changes — so the H2D traffic of a tick is O(changes), not O(N*S).
 * A delta rewrites one code of one record in place (bit-sliced fields: one bit per plane, by integer atomics).
 * Contract: within ONE bf_resident_apply call at most one delta per (run, field, index) — the host mirror
 * coalesces (the last value of a tick wins); deltas of different steps that share a word are fine.              */
enum {
  BF_DELTA_PHASE = 0,       /* index = step, code = BF_PHASE_* (0..14)                                           */
  BF_DELTA_COND = 1,        /* index = step, code = BF_COND_*                                                    */
  BF_DELTA_DECISION = 2,    /* index = step, code = BF_DEC_*                                                     */
  BF_DELTA_CHILD = 3,       /* index = child nibble (child_first[p] + branch), code = BF_PHASE_*                 */
  BF_DELTA_RUN_FLAGS = 4,   /* index ignored, code = BF_RF_* byte                                                */
  BF_DELTA_REGISTERED = 5,  /* index = parallel desc p, code = 0 / 1 (children_registered bit)                   */
  BF_DELTA_TOPO_SLOT = 6    /* index = low 16 bits, code = bits 16-23 of the new topology slot (slot < 2^24)     */
};
typedef struct bf_delta { /* 8 B */
  uint32_t run;
  uint16_t index;
  uint8_t field; /* BF_DELTA_* */
  uint8_t code;
} bf_delta;

/* A resident batch of up to `capacity` runs with the given layout; *handle_out identifies it on this ctx.       */
int bf_resident_create(bf_ctx* ctx, const bf_layout* layout, uint32_t capacity, uint32_t* handle_out);
int bf_resident_destroy(bf_ctx* ctx, uint32_t handle);
/* Full records for runs [first_run, first_run + n_runs) (new StoryRuns, or a resync): host buffer, synchronous.  */
int bf_resident_upload(bf_ctx* ctx, uint32_t handle, uint32_t first_run, uint32_t n_runs, const void* state_records);
/* Apply n deltas (host buffer; copied and scattered on the ctx stream, synchronous).                            */
int bf_resident_apply(bf_ctx* ctx, uint32_t handle, const bf_delta* deltas, uint32_t n);
/* One pass over runs [0, n_runs) of the resident state: kernels + D2H of the result records (host buffer) and
 * counts.  flags = BF_EVAL_* (expansion is not offered on this path).  Also the batch bf_schedule refers to.     */
int bf_resident_eval(bf_ctx* ctx, uint32_t handle, uint32_t n_runs, uint32_t flags, uint32_t max_iterations, void* result,
                     bf_counts* counts);
/* A whole tick in one call and one synchronisation: apply n_deltas deltas (may be 0), run the pass, read the results
 * (the kernels of later run chunks overlap the download of earlier ones).                                       */
int bf_resident_tick(bf_ctx* ctx, uint32_t handle, const bf_delta* deltas, uint32_t n_deltas, uint32_t n_runs, uint32_t flags,
                     uint32_t max_iterations, void* result, bf_counts* counts);
/* Read the device copy back (tests / resync checks).                                                            */
int bf_resident_download(bf_ctx* ctx, uint32_t handle, uint32_t first_run, uint32_t n_runs, void* state_records);

/* ------------------------------------------------------------------ compact results
 * What the consumer of a pass needs (findAndLaunchReadySteps, dag.go:1735-1775) is two short LISTS per run — the
 * ready steps to hand to StepExecutor.Execute and the skipped steps to mark — plus the run's summary flags, not 80+
 * bytes of masks per run.  The compact entry points turn the result masks into lists ON THE DEVICE and ship only
 * those:
 *   head[r]   one u32 per run: low 15 bits = low bits of the run's BF_SUM_* word (group, done / failed flags,
 *             PHASE_CHANGED; 0x7FFF = dead topology slot), bit 15 = the run's events are in this call's list,
 *             high 16 bits = how many.  The fixpoint iteration count of the summary word is not carried.
 *   events[]  one u16 per (run, step) that has any result bit set: step | kind << 10, run-major in batch order and
 *             step-ascending inside a run (the order of the reference's lists); run r's events are the next
 *             head[r] >> 16 entries.
 * At BASELINE configs[2] that is 0.4 MB + ~0.6 MB per pass instead of 8 MB of mask records.  The dense records stay
 * on the device (bf_schedule may follow).
 * BF_EVAL_CHANGED_ONLY (resident ticks): only runs whose result record differs from the PREVIOUS tick's carry
 * BF_HEAD_LISTED and events — the batcher keeps the last row of every run anyway, so a steady-state
 * tick returns O(changes) in both directions.  The first tick after bf_resident_create / _upload lists every run.   */
#define BF_EVT_READY 0x1u      /* step is in `ready`                                                            */
#define BF_EVT_SKIP 0x2u       /* step is in `skip`  (failed dependency or `if` false)                          */
#define BF_EVT_FAIL 0x4u       /* step is in `fail`        (only when the layout has BF_F_OUT_FAIL)              */
#define BF_EVT_NEEDS_COND 0x8u /* step is in `needs_cond`  (only with BF_F_OUT_NEEDS_COND)                       */
#define BF_EVT_SKIP_DEP 0x10u  /* step is in `skip_dep`    (only with BF_F_OUT_SKIP_DEP)                         */
#define BF_EVENT_STEP(e) ((uint32_t)(e) & 0x3FFu)
#define BF_EVENT_KIND(e) ((uint32_t)(e) >> 10)
#define BF_HEAD_SUMMARY_MASK 0x7FFFu
#define BF_HEAD_DEAD 0x7FFFu   /* low bits of a run whose topology slot is dead: no events                       */
#define BF_HEAD_LISTED 0x8000u
#define BF_HEAD_COUNT_SHIFT 16
#define BF_EVAL_CHANGED_ONLY 0x10u /* flag of bf_resident_tick_compact                                           */

typedef struct bf_compact_out {
  uint32_t struct_size;
  uint32_t n_listed;           /* out: runs that carry BF_HEAD_LISTED                                            */
  uint32_t* head;              /* [n_runs]                                                                       */
  uint16_t* events;            /* [events_cap]                                                                   */
  uint64_t events_cap;
  uint64_t n_events;           /* out: events the pass produced; when > events_cap only the first events_cap were written
                                  (heads always carry the true counts)                                           */
} bf_compact_out;

/* bf_eval with compact results: H2D(state) -> kernels -> D2H(heads + events).  batch->result is ignored
 * (may be NULL).  Host buffers, synchronous.                                                                     */
int bf_eval_compact(bf_ctx* ctx, const bf_batch* batch, bf_compact_out* out);
/* bf_resident_tick with compact results: deltas -> scatter -> pass -> heads + events.  This is the steady-state
 * tick of the operator: both directions are proportional to what changed.  Buffers in PINNED host memory
 * (bf_alloc_pinned, cudaHostAlloc, cudaHostRegister) are used in place: the scatter kernel reads the deltas from
 * them and the compaction kernel posts head words and events into them, with no staging copy on either side;
 * any other buffer goes through the upload / download copies.  Either way the call returns when they are filled. */
int bf_resident_tick_compact(bf_ctx* ctx, uint32_t handle, const bf_delta* deltas, uint32_t n_deltas, uint32_t n_runs,
                             uint32_t flags, uint32_t max_iterations, bf_compact_out* out, bf_counts* counts);

/* ------------------------------------------------------------------ device groups (SURVEY.md row e)
 * The reference is ONE operator process (cmd/main.go:220, StoryRunReconciler storyrun_controller.go:216), so the
 * multi-GPU path must be reachable from one process through this ABI: a group owns one ctx per device and an NCCL
 * communicator over them (ncclCommInitAll; libnccl.so.2 is loaded at bf_group_create, the library itself does not
 * link it).  StoryRuns are independent, so a batch shards into contiguous blocks of ceil(N / G) runs with NO
 * data-path collective; the one exchange of a pass is an ncclAllGather of each shard's 32-byte counts block, which
 * gives every shard the global totals and the offsets of its compacted lists.  Topologies live where their runs
 * are: put them through the shard's own ctx (bf_group_ctx + bf_topology_put_many; slot ids are per shard) or
 * replicate a shared set on every device with bf_group_topology_put_many (same slot ids everywhere).
 * Collective failures come back as BF_ENCCL with the NCCL text in bf_group_last_error.                            */
typedef struct bf_group bf_group;

int bf_group_create(bf_group** out, const int32_t* devices, uint32_t n_devices, const bf_config* cfg /* device field ignored; may be NULL */);
void bf_group_destroy(bf_group* g);
uint32_t bf_group_size(const bf_group* g);
bf_ctx* bf_group_ctx(bf_group* g, uint32_t shard);
const char* bf_group_last_error(const bf_group* g);
/* Runs [*first, *first + *count) of an n_runs batch belong to `shard`.                                            */
int bf_group_shard_range(const bf_group* g, uint32_t n_runs, uint32_t shard, uint32_t* first, uint32_t* count);
/* The same topologies on every device; fails with BF_ETOPO if the shards' slot tables have diverged (mixing
 * replicated and per-shard uploads is allowed only when the replicated ones come first).                          */
int bf_group_topology_put_many(bf_group* g, const bf_topology* topos, uint32_t count, uint32_t* slots_out);
/* One pass over a HOST batch sharded across the group: every shard runs bf_eval on its block of runs (its own H2D /
 * kernels / D2H pipeline, all shards concurrently), then the counts are all-gathered on the devices.
 * shard_counts[n_devices] (may be NULL) receives the gathered per-shard counts, batch->counts the global totals.
 * BF_EVAL_EXPANSION is not offered here (per-shard lists: call bf_eval on bf_group_ctx).                           */
int bf_group_eval(bf_group* g, const bf_batch* batch, bf_counts* shard_counts);
/* bf_schedule for a batch evaluated by the IMMEDIATELY PRECEDING bf_group_eval: every shard counts its own Running
 * StepRuns and demand, the per-story / per-queue / global totals are all-reduced (sum) and the per-queue highest
 * effective priority (max) across the shards, then every shard truncates its ready sets against the GLOBAL totals —
 * the limits and the priority ordering hold across the whole batch, as enforcePriorityOrdering / enforceSchedulingLimits
 * (dag.go:1801-1946) hold across the cluster.  runs[n_runs] and out->records cover the whole batch; the totals arrays
 * of `out` receive the global values.                                                                             */
int bf_group_schedule(bf_group* g, const bf_batch* batch, const bf_sched_run* runs, const bf_sched_tables* tables, bf_sched_out* out);

/* Pinned host memory for batches (cgo: memory with no Go pointers).           */
int bf_alloc_pinned(bf_ctx* ctx, size_t bytes, void** out);
int bf_free_pinned(bf_ctx* ctx, void* p);

/* Introspection used by bench/roofline accounting.                            */
typedef struct bf_stats {
  uint64_t kernel_launches;   /* kernels launched by this ctx so far            */
  uint64_t arena_used_bytes;  /* HBM bytes of topology records                  */
  uint64_t arena_cap_bytes;
  uint32_t n_topologies;
  uint32_t sm_count;
  uint32_t last_grid, last_block, last_smem_bytes, last_stages;
  uint32_t last_kernel;        /* 0 = one run per warp (general), 1 = packed lanes, 2 = packed lanes + deferred general */
  uint32_t last_runs_per_trip; /* StoryRuns evaluated per warp trip by the last pass                                   */
  uint32_t last_eval_chunks;   /* run chunks the last bf_eval pipelined over its copy/compute streams (1 = serial)      */
  uint32_t arena_compactions;  /* times the topology arena was re-packed (dropped records reclaimed)                             */
} bf_stats;
int bf_get_stats(const bf_ctx* ctx, bf_stats* out);
/* Device address + byte size of a topology record (for traffic accounting).   */
int bf_topology_record(const bf_ctx* ctx, uint32_t slot, uint64_t* dev_addr, uint32_t* bytes);
/* Host-only (no ctx, no GPU): validates `topo` exactly as bf_topology_put does and writes the device record it would upload
 * (the row format chosen for it, device_record.h) into out[cap]; *bytes_out = the record's size, also when cap is too
 * small (then nothing is written and BF_ENOMEM is returned).  For tests and for sizing an arena.                          */
int bf_topology_record_build(const bf_topology* topo, void* out, uint32_t cap, uint32_t* bytes_out);

#ifdef __cplusplus
}
#endif
#endif /* BOBRAFRONTIER_H_ */
