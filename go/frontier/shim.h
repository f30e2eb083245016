/*
 * shim.h — flat C wrappers over include/bobrafrontier.h for the cgo binding (go/frontier/frontier.go).
 *
 * cgo forbids passing a pointer to Go memory that itself contains Go pointers ("cgo argument has Go pointer to
 * unpinned Go pointer").  The ABI's argument structs (bf_topology, bf_batch, bf_sched_tables, bf_sched_out,
 * bf_compact_out) hold pointers, so they must never be built in Go memory.  Every wrapper below takes the scalars and
 * the buffer pointers as FLAT arguments — a Go caller passes &slice[0] of slices that hold no pointers, which is legal —
 * and builds the struct here, on the C stack.
 *
 * The same header is compiled by tests/c_abi_harness.c (gcc, no Go needed), which performs the call sequence of
 * frontier.go against the real library on a GPU: the struct layouts and the argument order the Go side relies on are
 * exercised even though this image has no Go toolchain.
 */
#ifndef BOBRAFRONTIER_GO_SHIM_H_
#define BOBRAFRONTIER_GO_SHIM_H_

#include <string.h>

#include "bobrafrontier.h"

static inline int bfgo_create(int32_t device, bf_ctx** out) {
  bf_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = (uint32_t)sizeof cfg;
  cfg.device = device;
  return bf_create(out, &cfg);
}

static inline int bfgo_topology_put(bf_ctx* c, uint32_t n_steps, uint32_t n_edges, const uint32_t* row_ptr, const uint16_t* col_idx,
                                    const uint8_t* step_flags, const bf_parallel_desc* parallel, uint32_t n_parallel,
                                    const uint8_t* allow_bits, uint32_t n_allow_bits, uint32_t* slot_out) {
  bf_topology t;
  memset(&t, 0, sizeof t);
  t.n_steps = n_steps; t.n_edges = n_edges; t.row_ptr = row_ptr; t.col_idx = col_idx; t.step_flags = step_flags;
  t.parallel = parallel; t.n_parallel = n_parallel; t.branch_allow_bits = allow_bits; t.n_branch_allow_bits = n_allow_bits;
  return bf_topology_put(c, &t, slot_out);
}

static inline void bfgo_batch(bf_batch* b, const bf_layout* L, uint32_t n_runs, uint32_t flags, uint32_t max_iterations, const void* state,
                              void* result, bf_counts* counts) {
  memset(b, 0, sizeof *b);
  b->struct_size = (uint32_t)sizeof *b;
  b->n_runs = n_runs; b->flags = flags; b->max_iterations = max_iterations; b->layout = *L;
  b->state = state; b->result = result; b->counts = counts;
}

static inline int bfgo_eval(bf_ctx* c, const bf_layout* L, uint32_t n_runs, uint32_t flags, uint32_t max_iterations, const void* state,
                            void* result, bf_counts* counts) {
  bf_batch b;
  bfgo_batch(&b, L, n_runs, flags, max_iterations, state, result, counts);
  return bf_eval(c, &b);
}

static inline void bfgo_compact(bf_compact_out* co, uint32_t* head, uint16_t* events, uint64_t events_cap) {
  memset(co, 0, sizeof *co);
  co->struct_size = (uint32_t)sizeof *co;
  co->head = head; co->events = events; co->events_cap = events_cap;
}

static inline int bfgo_eval_compact(bf_ctx* c, const bf_layout* L, uint32_t n_runs, uint32_t flags, uint32_t max_iterations,
                                    const void* state, uint32_t* head, uint16_t* events, uint64_t events_cap,
                                    uint64_t* n_events, uint32_t* n_listed, bf_counts* counts) {
  bf_batch b;
  bf_compact_out co;
  bfgo_batch(&b, L, n_runs, flags, max_iterations, state, NULL, counts);
  bfgo_compact(&co, head, events, events_cap);
  const int rc = bf_eval_compact(c, &b, &co);
  *n_events = co.n_events;
  *n_listed = co.n_listed;
  return rc;
}

static inline int bfgo_resident_tick_compact(bf_ctx* c, uint32_t handle, const bf_delta* deltas, uint32_t n_deltas, uint32_t n_runs,
                                             uint32_t flags, uint32_t max_iterations, uint32_t* head, uint16_t* events,
                                             uint64_t events_cap, uint64_t* n_events, uint32_t* n_listed, bf_counts* counts) {
  bf_compact_out co;
  bfgo_compact(&co, head, events, events_cap);
  const int rc = bf_resident_tick_compact(c, handle, deltas, n_deltas, n_runs, flags, max_iterations, &co, counts);
  *n_events = co.n_events;
  *n_listed = co.n_listed;
  return rc;
}

static inline void bfgo_tables(bf_sched_tables* t, uint32_t n_stories, uint32_t n_queues, int32_t global_limit, uint32_t global_base,
                               const int32_t* story_limit, const uint32_t* story_base, const int32_t* queue_limit,
                               const int32_t* queue_aging_s, const uint32_t* queue_base, const int32_t* queue_max_priority_base) {
  memset(t, 0, sizeof *t);
  t->struct_size = (uint32_t)sizeof *t;
  t->n_stories = n_stories; t->n_queues = n_queues; t->global_limit = global_limit; t->global_running_base = global_base;
  t->story_limit = story_limit; t->story_running_base = story_base; t->queue_limit = queue_limit; t->queue_aging_s = queue_aging_s;
  t->queue_running_base = queue_base; t->queue_max_priority_base = queue_max_priority_base;
}

static inline void bfgo_sched_out(bf_sched_out* o, void* records, uint32_t* story_running, uint32_t* queue_running,
                                  int32_t* queue_max_priority, uint32_t* global_running) {
  memset(o, 0, sizeof *o);
  o->struct_size = (uint32_t)sizeof *o;
  o->records = records; o->story_running = story_running; o->queue_running = queue_running;
  o->queue_max_priority = queue_max_priority; o->global_running = global_running;
}

static inline int bfgo_schedule(bf_ctx* c, const bf_layout* L, uint32_t n_runs, const bf_sched_run* runs, uint32_t n_stories,
                                uint32_t n_queues, int32_t global_limit, uint32_t global_base, const int32_t* story_limit,
                                const uint32_t* story_base, const int32_t* queue_limit, const int32_t* queue_aging_s,
                                const uint32_t* queue_base, const int32_t* queue_max_priority_base, void* records,
                                uint32_t* story_running, uint32_t* queue_running, int32_t* queue_max_priority, uint32_t* global_running) {
  bf_batch b;
  bf_sched_tables t;
  bf_sched_out o;
  bfgo_batch(&b, L, n_runs, 0, 0, NULL, NULL, NULL);
  bfgo_tables(&t, n_stories, n_queues, global_limit, global_base, story_limit, story_base, queue_limit, queue_aging_s, queue_base,
              queue_max_priority_base);
  bfgo_sched_out(&o, records, story_running, queue_running, queue_max_priority, global_running);
  return bf_schedule(c, &b, runs, &t, &o);
}

static inline int bfgo_group_eval(bf_group* g, const bf_layout* L, uint32_t n_runs, uint32_t flags, uint32_t max_iterations,
                                  const void* state, void* result, bf_counts* shard_counts, bf_counts* global_counts) {
  bf_batch b;
  bfgo_batch(&b, L, n_runs, flags, max_iterations, state, result, global_counts);
  return bf_group_eval(g, &b, shard_counts);
}

static inline int bfgo_group_schedule(bf_group* g, const bf_layout* L, uint32_t n_runs, const bf_sched_run* runs, uint32_t n_stories,
                                      uint32_t n_queues, int32_t global_limit, uint32_t global_base, const int32_t* story_limit,
                                      const uint32_t* story_base, const int32_t* queue_limit, const int32_t* queue_aging_s,
                                      const uint32_t* queue_base, const int32_t* queue_max_priority_base, void* records,
                                      uint32_t* story_running, uint32_t* queue_running, int32_t* queue_max_priority,
                                      uint32_t* global_running) {
  bf_batch b;
  bf_sched_tables t;
  bf_sched_out o;
  bfgo_batch(&b, L, n_runs, 0, 0, NULL, NULL, NULL);
  bfgo_tables(&t, n_stories, n_queues, global_limit, global_base, story_limit, story_base, queue_limit, queue_aging_s, queue_base,
              queue_max_priority_base);
  bfgo_sched_out(&o, records, story_running, queue_running, queue_max_priority, global_running);
  return bf_group_schedule(g, &b, runs, &t, &o);
}

#endif /* BOBRAFRONTIER_GO_SHIM_H_ */
