/*
 * c_abi_harness.c — plain C caller of libbobrafrontier.so that performs the call sequence of the cgo binding
 * (go/frontier/frontier.go) through the SAME flat wrappers (go/frontier/shim.h): New, PutTopology, Eval, EvalCompact,
 * Schedule, RedriveClosure, Resident.{Upload,Tick}, Group.{Eval,Schedule}.  The image has no Go toolchain; this is what
 * proves that the struct layouts, the argument order and the ownership rules the Go side relies on work against the real
 * library.  Expected answers are the config-1 lifecycle of SURVEY.md 8.2 (story A; B needs A; C needs B) and the
 * reference's own vectors (dag_test.go:842 skipped-by-failed-dependency, :744 concurrency limit).
 *
 *   gcc -O1 -Wall -Iinclude -Igo/frontier tests/c_abi_harness.c -Lbobrapet_b200/lib -lbobrafrontier -Wl,-rpath,... -o harness
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "shim.h"

#define CHECK(cond)                                                                 \
  do {                                                                              \
    if (!(cond)) { fprintf(stderr, "c_abi_harness: FAILED %s (line %d): %s\n", #cond, __LINE__, ctx ? bf_last_error(ctx) : ""); return 1; } \
  } while (0)

static void set_code(uint8_t* field, uint32_t words, int nbits, uint32_t i, uint32_t code) {
  uint32_t* w = (uint32_t*)field;
  for (int b = 0; b < nbits; ++b) {
    if ((code >> b) & 1u) w[(uint32_t)b * words + (i >> 5)] |= 1u << (i & 31u);
    else w[(uint32_t)b * words + (i >> 5)] &= ~(1u << (i & 31u));
  }
}
static uint32_t mask_of(const uint8_t* rec, uint32_t off) { return *(const uint32_t*)(rec + off); }

int main(void) {
  bf_ctx* ctx = NULL;
  CHECK(bf_abi_version() == BF_ABI_VERSION);
  CHECK(bfgo_create(0, &ctx) == BF_OK);

  /* ---- PutTopology: A; B needs A; C needs B (all engram steps, main group) ---- */
  const uint32_t row_ptr[4] = {0, 0, 1, 2};
  const uint16_t col_idx[2] = {0, 1};
  const uint8_t flags[3] = {BF_STEP_ENGRAM, BF_STEP_ENGRAM, BF_STEP_ENGRAM};
  uint32_t slot = 99;
  CHECK(bfgo_topology_put(ctx, 3, 2, row_ptr, col_idx, flags, NULL, 0, NULL, 0, &slot) == BF_OK);
  /* a cycle is rejected like validateRuntimeDependencyGraph does (dag_test.go:321) */
  const uint32_t rp_cyc[2] = {0, 1};
  const uint16_t ci_cyc[1] = {0};
  uint32_t bad = 0;
  CHECK(bfgo_topology_put(ctx, 1, 1, rp_cyc, ci_cyc, flags, NULL, 0, NULL, 0, &bad) == BF_ETOPO);

  /* ---- NewBatch: layout + pinned buffers ---- */
  bf_layout L;
  CHECK(bf_layout_init(&L, 3, 0, BF_F_OUT_SKIP_DEP) == BF_OK);
  enum { N = 6 };
  void *state_v = NULL, *result_v = NULL;
  CHECK(bf_alloc_pinned(ctx, (size_t)N * L.state_stride, &state_v) == BF_OK);
  CHECK(bf_alloc_pinned(ctx, (size_t)N * L.result_stride, &result_v) == BF_OK);
  uint8_t* state = (uint8_t*)state_v;
  uint8_t* result = (uint8_t*)result_v;
  memset(state, 0, (size_t)N * L.state_stride);
  /* runs: 0 nothing started (failFast) | 1 A Running | 2 A Succeeded | 3 A,B Succeeded | 4 A Failed, !failFast | 5 A Failed, failFast */
  const uint32_t phases[N][3] = {{0, 0, 0}, {BF_PHASE_RUNNING, 0, 0}, {BF_PHASE_SUCCEEDED, 0, 0}, {BF_PHASE_SUCCEEDED, BF_PHASE_SUCCEEDED, 0},
                                 {BF_PHASE_FAILED, 0, 0}, {BF_PHASE_FAILED, 0, 0}};
  const uint8_t rflags[N] = {BF_RF_FAIL_FAST, BF_RF_FAIL_FAST, BF_RF_FAIL_FAST, BF_RF_FAIL_FAST, 0, BF_RF_FAIL_FAST};
  for (uint32_t r = 0; r < N; ++r) {
    uint8_t* rec = state + (size_t)r * L.state_stride;
    bf_run_header* h = (bf_run_header*)rec;
    h->topo_slot = slot;
    h->run_flags = rflags[r];
    for (uint32_t i = 0; i < 3; ++i) set_code(rec + L.off_phase, L.words, 4, i, phases[r][i]);
  }

  /* ---- Eval ---- */
  bf_counts counts;
  CHECK(bfgo_eval(ctx, &L, N, BF_EVAL_VALIDATE, 0, state, result, &counts) == BF_OK);
  const uint32_t want_ready[N] = {1u, 0u, 2u, 4u, 0u, 0u};   /* {A} {} {B} {C} {} {} */
  const uint32_t want_skip[N] = {0u, 0u, 0u, 0u, 2u, 0u};    /* run 4: B skipped (failed dependency A), dag_test.go:842 */
  for (uint32_t r = 0; r < N; ++r) {
    const uint8_t* rec = result + (size_t)r * L.result_stride;
    CHECK(mask_of(rec, L.off_ready) == want_ready[r]);
    CHECK(mask_of(rec, L.off_skip) == want_skip[r]);
  }
  CHECK(mask_of(result + 4 * (size_t)L.result_stride, L.off_skip_dep) == 2u);
  /* run 5: fail-fast marks B and C Skipped => main done, group = finalize on host (dag.go:3289-3312, 490-495) */
  const bf_result_header* h5 = (const bf_result_header*)(result + 5 * (size_t)L.result_stride);
  CHECK((h5->summary & BF_SUM_GROUP_MASK) == BF_GROUP_DONE && (h5->summary & BF_SUM_MAIN_FAILED) && (h5->summary & BF_SUM_PHASE_CHANGED));
  CHECK(counts.ready == 3 && counts.skip == 1 && counts.evals == 3 * N);

  /* ---- EvalCompact: the same pass as lists (head word per run + 16-bit events) ---- */
  uint32_t head[N], n_listed = 0;
  uint16_t events[16];
  uint64_t n_events = 0;
  bf_counts c2;
  CHECK(bfgo_eval_compact(ctx, &L, N, 0, 0, state, head, events, 16, &n_events, &n_listed, &c2) == BF_OK);
  CHECK(n_events == 4 && n_listed == N && memcmp(&c2, &counts, sizeof counts) == 0);
  const uint32_t want_count[N] = {1, 0, 1, 1, 1, 0};
  for (uint32_t r = 0; r < N; ++r) CHECK((head[r] >> BF_HEAD_COUNT_SHIFT) == want_count[r] && (head[r] & BF_HEAD_LISTED));
  CHECK(BF_EVENT_STEP(events[0]) == 0 && BF_EVENT_KIND(events[0]) == BF_EVT_READY);                        /* run 0: A ready */
  CHECK(BF_EVENT_STEP(events[1]) == 1 && BF_EVENT_KIND(events[1]) == BF_EVT_READY);                        /* run 2: B ready */
  CHECK(BF_EVENT_STEP(events[2]) == 2 && BF_EVENT_KIND(events[2]) == BF_EVT_READY);                        /* run 3: C ready */
  CHECK(BF_EVENT_STEP(events[3]) == 1 && BF_EVENT_KIND(events[3]) == (BF_EVT_SKIP | BF_EVT_SKIP_DEP));     /* run 4: B skipped */
  CHECK((head[5] & BF_HEAD_SUMMARY_MASK) == (h5->summary & BF_HEAD_SUMMARY_MASK));

  /* ---- Schedule: story 0 has limit 2 and already 1 Running StepRun in the batch (run 1's A) => one slot for three ready steps */
  CHECK(bfgo_eval(ctx, &L, N, 0, 0, state, result, &counts) == BF_OK);
  bf_sched_run runs[N];
  memset(runs, 0, sizeof runs);
  for (uint32_t r = 0; r < N; ++r) { runs[r].queued_elapsed_s = BF_SCHED_NONE; runs[r].run_phase = BF_PHASE_RUNNING; }
  const int32_t story_limit[1] = {2}, queue_limit[1] = {0}, queue_aging[1] = {0};
  uint8_t sched[N * BF_SCHED_STRIDE(1)];
  uint32_t story_running[1], queue_running[1], global_running[1];
  int32_t queue_maxprio[1];
  CHECK(bfgo_schedule(ctx, &L, N, runs, 1, 1, 0, 0, story_limit, NULL, queue_limit, queue_aging, NULL, NULL, sched, story_running,
                      queue_running, queue_maxprio, global_running) == BF_OK);
  CHECK(story_running[0] == 1 && global_running[0] == 1);
  uint32_t launched = 0, queued = 0;
  for (uint32_t r = 0; r < N; ++r) {
    const bf_sched_header* sh = (const bf_sched_header*)(sched + (size_t)r * BF_SCHED_STRIDE(1));
    launched += sh->n_launch;
    queued += sh->n_queued_story;
  }
  /* every run is limited independently against the same snapshot (StepRuns created in this tick do not count yet):
     each of the three ready runs sees 2 - 1 = 1 slot and launches its single ready step */
  CHECK(launched == 3 && queued == 0);
  const int32_t tight[1] = {1};
  CHECK(bfgo_schedule(ctx, &L, N, runs, 1, 1, 0, 0, tight, NULL, queue_limit, queue_aging, NULL, NULL, sched, story_running, queue_running,
                      queue_maxprio, global_running) == BF_OK);
  launched = queued = 0;
  for (uint32_t r = 0; r < N; ++r) {
    const bf_sched_header* sh = (const bf_sched_header*)(sched + (size_t)r * BF_SCHED_STRIDE(1));
    launched += sh->n_launch;
    queued += sh->n_queued_story;
  }
  CHECK(launched == 0 && queued == 3);   /* limit 1, 1 running: ready 0, the rest queued (dag_test.go:744) */

  /* ---- RedriveClosure: a redrive from B resets B and C (storyrun_controller.go:535-558) ---- */
  const uint32_t q_slot[1] = {slot}, q_step[1] = {1};
  uint32_t closure[1] = {0};
  CHECK(bf_topology_closure(ctx, q_slot, q_step, 1, 1, closure) == BF_OK);
  CHECK(closure[0] == 6u);

  /* ---- Resident: full records once, then a delta tick (A of run 0 -> Succeeded => B ready) ---- */
  uint32_t handle = 0;
  CHECK(bf_resident_create(ctx, &L, N, &handle) == BF_OK);
  CHECK(bf_resident_upload(ctx, handle, 0, N, state) == BF_OK);
  bf_delta d;
  d.run = 0; d.index = 0; d.field = BF_DELTA_PHASE; d.code = BF_PHASE_SUCCEEDED;
  CHECK(bfgo_resident_tick_compact(ctx, handle, &d, 1, N, 0, 0, head, events, 16, &n_events, &n_listed, &c2) == BF_OK);
  CHECK(n_events == 4 && n_listed == N && BF_EVENT_STEP(events[0]) == 1 && BF_EVENT_KIND(events[0]) == BF_EVT_READY);
  /* a changed-only tick with one more delta (B of run 0 -> Succeeded => C ready): only run 0 is listed */
  d.index = 1;
  CHECK(bfgo_resident_tick_compact(ctx, handle, &d, 1, N, BF_EVAL_CHANGED_ONLY, 0, head, events, 16, &n_events, &n_listed, &c2) == BF_OK);
  CHECK(n_listed == 1 && n_events == 1 && (head[0] & BF_HEAD_LISTED) && !(head[2] & BF_HEAD_LISTED) && BF_EVENT_STEP(events[0]) == 2);
  CHECK(bf_resident_destroy(ctx, handle) == BF_OK);

  /* ---- error behaviour: a bad struct size is BF_EINVAL with text, nothing aborts ---- */
  bf_batch bb;
  memset(&bb, 0, sizeof bb);
  CHECK(bf_eval(ctx, &bb) == BF_EINVAL && strlen(bf_last_error(ctx)) > 0);

  CHECK(bf_free_pinned(ctx, state_v) == BF_OK && bf_free_pinned(ctx, result_v) == BF_OK);
  bf_destroy(ctx);
  ctx = NULL;

  /* ---- Group: one process, every device it is given (here: device 0), counts gathered through NCCL ---- */
  bf_group* g = NULL;
  const int32_t devs[1] = {0};
  CHECK(bf_group_create(&g, devs, 1, NULL) == BF_OK);
  CHECK(bf_group_size(g) == 1);
  uint32_t first = 9, cnt = 9;
  CHECK(bf_group_shard_range(g, N, 0, &first, &cnt) == BF_OK && first == 0 && cnt == N);
  bf_topology t;
  memset(&t, 0, sizeof t);
  t.n_steps = 3; t.n_edges = 2; t.row_ptr = row_ptr; t.col_idx = col_idx; t.step_flags = flags;
  uint32_t gslot = 99;
  CHECK(bf_group_topology_put_many(g, &t, 1, &gslot) == BF_OK);
  uint8_t* gstate = (uint8_t*)calloc(N, L.state_stride);
  uint8_t* gresult = (uint8_t*)calloc(N, L.result_stride);
  for (uint32_t r = 0; r < N; ++r) {
    uint8_t* rec = gstate + (size_t)r * L.state_stride;
    ((bf_run_header*)rec)->topo_slot = gslot;
    ((bf_run_header*)rec)->run_flags = rflags[r];
    for (uint32_t i = 0; i < 3; ++i) set_code(rec + L.off_phase, L.words, 4, i, phases[r][i]);
  }
  bf_counts per[1], glob;
  if (bfgo_group_eval(g, &L, N, 0, 0, gstate, gresult, per, &glob) != BF_OK) { fprintf(stderr, "c_abi_harness: group eval: %s\n", bf_group_last_error(g)); return 1; }
  if (!(glob.ready == 3 && glob.skip == 1 && per[0].evals == 3 * N)) { fprintf(stderr, "c_abi_harness: group counts differ\n"); return 1; }
  for (uint32_t r = 0; r < N; ++r)
    if (mask_of(gresult + (size_t)r * L.result_stride, L.off_ready) != want_ready[r]) { fprintf(stderr, "c_abi_harness: group result differs\n"); return 1; }
  if (bfgo_group_schedule(g, &L, N, runs, 1, 1, 0, 0, tight, NULL, queue_limit, queue_aging, NULL, NULL, sched, story_running, queue_running,
                          queue_maxprio, global_running) != BF_OK) { fprintf(stderr, "c_abi_harness: group schedule: %s\n", bf_group_last_error(g)); return 1; }
  if (story_running[0] != 1) { fprintf(stderr, "c_abi_harness: group schedule totals differ\n"); return 1; }
  free(gstate);
  free(gresult);
  bf_group_destroy(g);
  printf("c_abi_harness: ok\n");
  return 0;
}
