/*
 * packed_ref.c — scalar CPU statement of the PACKED frontier contract.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): may be called from tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * never from the product path.
 *
 * It evaluates exactly the records the CUDA kernel sees (include/bobrafrontier.h)
 * with plain per-step loops, one StoryRun at a time, following the reference
 * (paths relative to /root/reference/internal/controller/runs):
 *
 *   stage G  checkSyncGates / Sleep / Wait           dag.go:1455-1547, 1217-1288, 1291-1452
 *   stage H  checkSyncParallelSteps                  dag.go:1112-1200
 *   stage I  fail-fast, topology-terminated,         dag.go:422-511, 3282-3342
 *            compensation skipping, group selection
 *   stage B  buildStateMaps + queued clearing        dag.go:3358-3391, 2020-2051
 *   stage D  findReadySteps                          dag.go:2631-2848
 *   stage J  runDagIterations fixpoint + launch      dag.go:393-540, 1735-1775;
 *            effects of Execute                      step_executor.go:132-185, 740-811
 *
 * It is validated against the object-shaped restatement (oracle/pyoracle.py,
 * oracle/refshape.cc), which in turn is pinned by the reference's own
 * known-answer tests (tests/test_oracle_kat.py).
 *
 * Build: gcc -O2 -shared -fPIC -pthread -I../include packed_ref.c -o _build/libpacked_ref.so
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "bobrafrontier.h"

typedef struct orc_topology {
  uint32_t n_steps, n_edges;
  const uint32_t* row_ptr;
  const uint16_t* col_idx;
  const uint8_t* step_flags;
  const bf_parallel_desc* parallel;
  uint32_t n_parallel;
  const uint8_t* branch_allow_bits;
  const uint32_t* child_first; /* [n_parallel] nibble offset of each desc's children in the child area */
} orc_topology;

static inline int lut(unsigned table, int p) { return (int)((table >> p) & 1u); }
static inline int is_term(int p) { return lut(BF_LUT_TERMINAL, p); }

static inline int get_nib(const uint8_t* a, uint32_t i) { return (a[i >> 1] >> ((i & 1u) * 4u)) & 0xF; }
/* bit-sliced code access: plane b at base + b*W words, step i = bit i%32 of word i/32 */
static inline int get_code(const uint8_t* base, uint32_t W, int nbits, uint32_t i) {
  const uint32_t* w = (const uint32_t*)base;
  int v = 0;
  for (int b = 0; b < nbits; ++b) v |= (int)((w[(uint32_t)b * W + (i >> 5)] >> (i & 31u)) & 1u) << b;
  return v;
}
static inline void set_code(uint8_t* base, uint32_t W, int nbits, uint32_t i, int v) {
  uint32_t* w = (uint32_t*)base;
  for (int b = 0; b < nbits; ++b)
    if ((v >> b) & 1) w[(uint32_t)b * W + (i >> 5)] |= 1u << (i & 31u);
}
static inline int get_bit(const uint8_t* a, uint32_t i) { return (a[i >> 3] >> (i & 7u)) & 1; }
static inline void set_mask(uint32_t* m, uint32_t i) { m[i >> 5] |= 1u << (i & 31u); }
static inline int tst_mask(const uint32_t* m, uint32_t i) { return (int)((m[i >> 5] >> (i & 31u)) & 1u); }

typedef struct run_scratch {
  uint8_t phase[BF_MAX_STEPS];
  uint8_t fail_now[BF_MAX_STEPS];
  uint32_t ready[BF_MAX_STEPS / 32], skip[BF_MAX_STEPS / 32], fail[BF_MAX_STEPS / 32];
  uint32_t needs_cond[BF_MAX_STEPS / 32], skip_dep[BF_MAX_STEPS / 32];
  uint32_t it_ready[BF_MAX_STEPS / 32], it_skip[BF_MAX_STEPS / 32];
} run_scratch;

/* dag.go:3377-3388 over the *current* phase, with the step's allowFailure bit */
static inline int st_completed(int p, uint8_t f) {
  return lut(BF_LUT_COMPLETED0, p) || (is_term(p) && (f & BF_SF_ALLOW_FAILURE));
}
static inline int st_failed(int p, uint8_t f) { return is_term(p) && !st_completed(p, f); }

static void eval_run(const orc_topology* T, const bf_layout* L, const uint8_t* srec, uint8_t* rrec,
                     uint32_t eflags, uint32_t max_iter, run_scratch* s, bf_counts* cnt) {
  const bf_run_header* rh = (const bf_run_header*)srec;
  const uint32_t S = T->n_steps;
  const uint32_t W = L->words;
  const uint8_t rf = rh->run_flags;
  const int fail_fast = (rf & BF_RF_FAIL_FAST) != 0;
  const int realtime = (rf & BF_RF_REALTIME) != 0;
  const int topo_term = (rf & BF_RF_TOPOLOGY_TERMINATED) != 0;
  const uint8_t* in_phase = srec + L->off_phase;
  const uint8_t* cond = (L->off_cond != BF_OFF_NONE) ? srec + L->off_cond : NULL;
  const uint8_t* dec = (L->off_decision != BF_OFF_NONE) ? srec + L->off_decision : NULL;
  const uint8_t* child = (L->off_child != BF_OFF_NONE) ? srec + L->off_child : NULL;
  const uint64_t registered = rh->children_registered;
  uint32_t nmain = 0, ncomp = 0, nfinal = 0;
  int changed = 0;

  for (uint32_t i = 0; i < S; ++i) {
    int p = get_code(in_phase, W, 4, i);
    if (p == BF_PHASE_RESERVED) p = BF_PHASE_NONE;
    s->phase[i] = (uint8_t)p;
    s->fail_now[i] = 0;
    int g = (T->step_flags[i] & BF_SF_GROUP_MASK) >> BF_SF_GROUP_SHIFT;
    if (g == BF_GROUP_MAIN) nmain++; else if (g == BF_GROUP_COMPENSATION) ncomp++; else nfinal++;
  }
  memset(s->ready, 0, sizeof s->ready); memset(s->skip, 0, sizeof s->skip);
  memset(s->fail, 0, sizeof s->fail); memset(s->needs_cond, 0, sizeof s->needs_cond);
  memset(s->skip_dep, 0, sizeof s->skip_dep);

  uint32_t summary = 0, n_exp = 0, iters = 0;
  uint32_t cap = (eflags & BF_EVAL_FIXPOINT) ? (max_iter ? max_iter : S + 1) : 1;

  for (uint32_t it = 0; it < cap; ++it) {
    iters++;
    /* ---- G: primitive syncs (gate/sleep/wait share one decision mapping) ---- */
    if (dec) {
      for (uint32_t i = 0; i < S; ++i) {
        uint8_t f = T->step_flags[i];
        int ty = f & BF_SF_TYPE_MASK;
        if (ty != BF_STEP_GATE && ty != BF_STEP_SLEEP && ty != BF_STEP_WAIT) continue;
        int p = s->phase[i];
        if (!lut(BF_LUT_RUNNING, p)) continue; /* exists, non-terminal, Paused|Running|Pending: dag.go:1469-1475 */
        int np;
        switch (get_code(dec, W, 2, i)) {
          case BF_DEC_SUCCEED: np = BF_PHASE_SUCCEEDED; break;                         /* :1502, :1253, :1412 */
          case BF_DEC_FAIL: np = BF_PHASE_FAILED; break;                               /* :1508 */
          case BF_DEC_TIMED_OUT: np = (f & BF_SF_ON_TIMEOUT_SKIP) ? BF_PHASE_SKIPPED : BF_PHASE_TIMEOUT; break; /* :1655-1668 */
          default: np = BF_PHASE_PAUSED; break;                                        /* :1515, :1270, :1416 */
        }
        if (np != p) { s->phase[i] = (uint8_t)np; changed = 1; }
      }
    }
    /* ---- H: parallel join, dag.go:1131-1198 ---- */
    if (child) {
      for (uint32_t q = 0; q < T->n_parallel; ++q) {
        const bf_parallel_desc* d = &T->parallel[q];
        int p = s->phase[d->step];
        if (p == BF_PHASE_NONE || is_term(p)) continue;         /* :1136-1139 */
        if (!((registered >> q) & 1ull) || d->branches == 0) continue; /* :1140-1143: no registered children */
        int all_done = 1, any_failed = 0;
        for (uint32_t b = 0; b < d->branches; ++b) {
          int cp = get_nib(child, T->child_first[q] + b);
          if (cp == BF_PHASE_RESERVED) cp = BF_PHASE_NONE;
          if (cp == BF_PHASE_NONE || !is_term(cp)) { all_done = 0; continue; }   /* :1165 */
          if (cp == BF_PHASE_SUCCEEDED || cp == BF_PHASE_SKIPPED) continue;      /* :1169 */
          if (T->branch_allow_bits && get_bit(T->branch_allow_bits, d->allow_first + b)) continue; /* :1172 */
          any_failed = 1;
        }
        if (!all_done) continue;
        s->phase[d->step] = (uint8_t)(any_failed ? BF_PHASE_FAILED : BF_PHASE_SUCCEEDED);
        changed = 1;
      }
    }
    /* ---- I: group selection, dag.go:422-495 ---- */
    int group;
    uint32_t sum = 0;
    if (rf & BF_RF_HOST_GROUP) {
      group = (rf >> BF_RF_HOST_GROUP_SHIFT) & 3;
    } else {
      int any_main_failed = 0;
      for (uint32_t i = 0; i < S; ++i) {
        uint8_t f = T->step_flags[i];
        if ((f & BF_SF_GROUP_MASK) != (BF_GROUP_MAIN << BF_SF_GROUP_SHIFT)) continue;
        if (st_failed(s->phase[i], f)) any_main_failed = 1;
      }
      if (fail_fast && any_main_failed) { /* markFailFastSkipped :3289-3312 */
        for (uint32_t i = 0; i < S; ++i) {
          uint8_t f = T->step_flags[i];
          if ((f & BF_SF_GROUP_MASK) != (BF_GROUP_MAIN << BF_SF_GROUP_SHIFT)) continue;
          int p = s->phase[i];
          if (st_completed(p, f) || lut(BF_LUT_RUNNING_Q, p)) continue;
          if (is_term(p)) continue;
          s->phase[i] = BF_PHASE_SKIPPED; changed = 1;
        }
      }
      uint32_t done_cnt = 0;
      for (uint32_t i = 0; i < S; ++i) {
        uint8_t f = T->step_flags[i];
        if ((f & BF_SF_GROUP_MASK) != (BF_GROUP_MAIN << BF_SF_GROUP_SHIFT)) continue;
        int p = s->phase[i];
        if (st_completed(p, f) || st_failed(p, f)) done_cnt++;
      }
      int main_done = (nmain == 0) || (done_cnt == nmain); /* stepsTerminal :3282 */
      if (!main_done && realtime && topo_term) {           /* :436-464 */
        main_done = 1;
        any_main_failed = 0;
        for (uint32_t i = 0; i < S; ++i) {
          uint8_t f = T->step_flags[i];
          if ((f & BF_SF_GROUP_MASK) != (BF_GROUP_MAIN << BF_SF_GROUP_SHIFT)) continue;
          int p = s->phase[i];
          if (p != BF_PHASE_NONE && !is_term(p)) { s->phase[i] = BF_PHASE_FAILED; changed = 1; }
          if (st_failed(s->phase[i], f)) any_main_failed = 1;
        }
      }
      if (main_done && !any_main_failed && ncomp > 0) { /* markCompensationsSkipped :3314-3342 */
        for (uint32_t i = 0; i < S; ++i) {
          uint8_t f = T->step_flags[i];
          if ((f & BF_SF_GROUP_MASK) != (BF_GROUP_COMPENSATION << BF_SF_GROUP_SHIFT)) continue;
          int p = s->phase[i];
          if (st_completed(p, f) || lut(BF_LUT_RUNNING, p) || st_failed(p, f)) continue;
          if (is_term(p)) continue;
          s->phase[i] = BF_PHASE_SKIPPED; changed = 1;
        }
      }
      uint32_t cdone = 0, fdone = 0;
      int any_comp_failed = 0, any_final_failed = 0;
      for (uint32_t i = 0; i < S; ++i) {
        uint8_t f = T->step_flags[i];
        int g = (f & BF_SF_GROUP_MASK) >> BF_SF_GROUP_SHIFT;
        int p = s->phase[i];
        int c = st_completed(p, f), fl = st_failed(p, f);
        if (g == BF_GROUP_COMPENSATION) { cdone += (uint32_t)(c || fl); any_comp_failed |= fl; }
        else if (g == BF_GROUP_FINALLY) { fdone += (uint32_t)(c || fl); any_final_failed |= fl; }
      }
      int comp_done = (ncomp == 0) || (cdone == ncomp);
      int final_done = (nfinal == 0) || (fdone == nfinal);
      if (!main_done) group = BF_GROUP_MAIN;                                          /* :484 */
      else if (any_main_failed && ncomp > 0 && !comp_done) group = BF_GROUP_COMPENSATION; /* :486 */
      else if (nfinal > 0 && !final_done) group = BF_GROUP_FINALLY;                   /* :488 */
      else group = BF_GROUP_DONE;                                                     /* :490 */
      if (main_done) sum |= BF_SUM_MAIN_DONE;
      if (any_main_failed) sum |= BF_SUM_MAIN_FAILED;
      if (comp_done) sum |= BF_SUM_COMP_DONE;
      if (final_done) sum |= BF_SUM_FINAL_DONE;
      if (any_comp_failed) sum |= BF_SUM_COMP_FAILED;
      if (any_final_failed) sum |= BF_SUM_FINAL_FAILED;
    }
    summary = sum | (uint32_t)group;

    memset(s->it_ready, 0, sizeof s->it_ready);
    memset(s->it_skip, 0, sizeof s->it_skip);
    int any_ready = 0, any_skip = 0, stop_ready = 0;
    if (group != BF_GROUP_DONE) {
      /* ---- D: findReadySteps, dag.go:2647-2846, in list order ---- */
      const int allow_failed = group != BF_GROUP_MAIN;                   /* :500 */
      const int skip_on_failed = group == BF_GROUP_MAIN && !fail_fast;   /* :501 */
      /* `completed`/`running` are built once (:497) from the pre-D snapshot; FAILs made
         inside the loop (:2744,:2810) change stepStates but not those maps. */
      static __thread uint8_t snap[BF_MAX_STEPS];
      memcpy(snap, s->phase, S);
      for (uint32_t i = 0; i < S; ++i) {
        uint8_t f = T->step_flags[i];
        if (((f & BF_SF_GROUP_MASK) >> BF_SF_GROUP_SHIFT) != group) continue;
        int p = snap[i];
        if (st_completed(p, f) || lut(BF_LUT_RUNNING_Q, p)) continue;   /* :2649 */
        if (is_term(p)) continue;                                       /* :2652-2708 */
        int any_unmet = 0, any_fd = 0;
        for (uint32_t e = T->row_ptr[i]; e < T->row_ptr[i + 1]; ++e) {
          uint32_t d = T->col_idx[e];
          int pd = snap[d];
          uint8_t fd = T->step_flags[d];
          int sat, fdep;
          if (s->fail_now[d] && d < i) {
            /* d was set Failed earlier in this same loop: not in `completed`, depState terminal */
            sat = allow_failed;               /* :2722 (Failed is terminal, never Succeeded for realtime :3465) */
            fdep = !sat && skip_on_failed;    /* :2725 */
          } else {
            sat = st_completed(pd, fd)                                  /* :2715 */
                  || (realtime && lut(BF_LUT_RT_SAT, pd))               /* :2719 */
                  || (allow_failed && is_term(pd));                     /* :2722 */
            fdep = !sat && skip_on_failed && is_term(pd);               /* :2725-2727 */
          }
          if (!sat) { any_unmet = 1; if (fdep) any_fd = 1; }
        }
        if (any_fd) { /* skip_max: SURVEY 8.0-F */
          set_mask(s->it_skip, i); set_mask(s->skip_dep, i); any_skip = 1;
          continue;
        }
        if (any_unmet) continue;
        if ((f & BF_SF_HAS_IF) && !realtime) set_mask(s->needs_cond, i);
        int c = cond ? get_code(cond, W, 2, i) : BF_COND_PASS;
        if (c == BF_COND_PASS) {
          set_mask(s->it_ready, i); any_ready = 1;
          if ((f & BF_SF_TYPE_MASK) == BF_STEP_STOP) stop_ready = 1;
        } else if (c == BF_COND_SKIP) {
          set_mask(s->it_skip, i); any_skip = 1;
        } else if (c == BF_COND_FAIL) {
          set_mask(s->fail, i); s->fail_now[i] = 1; s->phase[i] = BF_PHASE_FAILED; changed = 1;
        }
      }
    }
    for (uint32_t w = 0; w < W; ++w) { s->ready[w] |= s->it_ready[w]; s->skip[w] |= s->it_skip[w]; }
    /* expansion count of this iteration's ready parallel steps */
    for (uint32_t q = 0; q < T->n_parallel; ++q)
      if (tst_mask(s->it_ready, T->parallel[q].step)) n_exp += T->parallel[q].branches;

    if (!(eflags & BF_EVAL_FIXPOINT)) break;
    if (group == BF_GROUP_DONE) break;                 /* finalize on host, :490-494 */
    /* ---- launch effects: dag.go:1735-1775, step_executor.go:132-185 ---- */
    for (uint32_t i = 0; i < S; ++i) {
      s->fail_now[i] = 0; /* Failed is now part of the snapshot */
      if (tst_mask(s->it_skip, i)) { s->phase[i] = BF_PHASE_SKIPPED; changed = 1; continue; }
      if (!tst_mask(s->it_ready, i)) continue;
      int ty = T->step_flags[i] & BF_SF_TYPE_MASK;
      int p = s->phase[i], np = p;
      switch (ty) {
        case BF_STEP_CONDITION: np = BF_PHASE_SUCCEEDED; break;                      /* :168-170 */
        case BF_STEP_SLEEP: case BF_STEP_GATE: case BF_STEP_WAIT: np = BF_PHASE_PAUSED; break; /* :171-179 */
        case BF_STEP_PARALLEL: np = BF_PHASE_RUNNING; break;                         /* :808-809 */
        case BF_STEP_STOP: break;                                                    /* host executes (with.phase) */
        default: if (p == BF_PHASE_NONE || p == BF_PHASE_PENDING_QUEUED) np = BF_PHASE_RUNNING; break; /* dag.go:1770-1774 */
      }
      /* a parallel step expanded in this pass has freshly created children with no phase yet
         (step_executor.go:792-806): its join cannot fire before the host reports them, so H
         keeps consulting only the children_registered bits the host supplied. */
      if (np != p) { s->phase[i] = (uint8_t)np; changed = 1; }
    }
    if (!any_ready && !any_skip) break;                /* :537 */
    if (stop_ready) break;                             /* contract: a ready `stop` hands the run to the host */
  }

  /* ---- write the result record ---- */
  bf_result_header* oh = (bf_result_header*)rrec;
  uint32_t nr = 0, ns = 0;
  for (uint32_t w = 0; w < W; ++w) { nr += (uint32_t)__builtin_popcount(s->ready[w]); ns += (uint32_t)__builtin_popcount(s->skip[w]); }
  if (changed) summary |= BF_SUM_PHASE_CHANGED;
  summary |= iters << BF_SUM_ITER_SHIFT;
  oh->summary = summary; oh->n_ready = nr; oh->n_skip = ns; oh->n_expansion = n_exp;
  memcpy(rrec + L->off_ready, s->ready, W * 4);
  memcpy(rrec + L->off_skip, s->skip, W * 4);
  if (L->off_fail != BF_OFF_NONE) memcpy(rrec + L->off_fail, s->fail, W * 4);
  if (L->off_needs_cond != BF_OFF_NONE) memcpy(rrec + L->off_needs_cond, s->needs_cond, W * 4);
  if (L->off_skip_dep != BF_OFF_NONE) memcpy(rrec + L->off_skip_dep, s->skip_dep, W * 4);
  if (L->off_phase_out != BF_OFF_NONE) {
    uint8_t* po = rrec + L->off_phase_out;
    memset(po, 0, W * 16);
    for (uint32_t i = 0; i < S; ++i) set_code(po, W, 4, i, s->phase[i]);
  }
  if (cnt) { cnt->ready += nr; cnt->skip += ns; cnt->expansion += n_exp; cnt->evals += S; }
}

typedef struct job {
  const orc_topology* topos; uint32_t n_topos;
  const bf_layout* L; const uint8_t* state; uint8_t* result;
  uint32_t lo, hi, eflags, max_iter; bf_counts cnt; int rc;
} job;

static void* worker(void* arg) {
  job* j = (job*)arg;
  run_scratch* s = (run_scratch*)malloc(sizeof(run_scratch));
  if (!s) { j->rc = BF_ENOMEM; return NULL; }
  memset(&j->cnt, 0, sizeof j->cnt);
  for (uint32_t r = j->lo; r < j->hi; ++r) {
    const uint8_t* srec = j->state + (size_t)r * j->L->state_stride;
    const bf_run_header* rh = (const bf_run_header*)srec;
    if (rh->topo_slot >= j->n_topos || j->topos[rh->topo_slot].n_steps == 0 ||
        j->topos[rh->topo_slot].n_steps > j->L->steps_max) { j->rc = BF_ETOPO; break; }
    eval_run(&j->topos[rh->topo_slot], j->L, srec, j->result + (size_t)r * j->L->result_stride,
             j->eflags, j->max_iter, s, &j->cnt);
  }
  free(s);
  return NULL;
}

/* Evaluate a batch.  topos is indexed by bf_run_header.topo_slot. */
int orc_packed_eval(const orc_topology* topos, uint32_t n_topos, const bf_layout* L, uint32_t n_runs,
                    const uint8_t* state, uint8_t* result, uint32_t eflags, uint32_t max_iter,
                    bf_counts* counts, int threads) {
  if (!topos || !L || (!state && n_runs) || (!result && n_runs)) return BF_EINVAL;
  if (L->steps_max == 0 || L->steps_max > BF_MAX_STEPS) return BF_EINVAL;
  if (threads < 1) threads = 1;
  if ((uint32_t)threads > n_runs) threads = n_runs ? (int)n_runs : 1;
  job* jobs = (job*)calloc((size_t)threads, sizeof(job));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  if (!jobs || !th) { free(jobs); free(th); return BF_ENOMEM; }
  uint32_t per = (n_runs + (uint32_t)threads - 1) / (uint32_t)threads;
  for (int t = 0; t < threads; ++t) {
    job* j = &jobs[t];
    j->topos = topos; j->n_topos = n_topos; j->L = L; j->state = state; j->result = result;
    j->lo = (uint32_t)t * per; j->hi = j->lo + per > n_runs ? n_runs : j->lo + per;
    if (j->lo > n_runs) j->lo = n_runs;
    j->eflags = eflags; j->max_iter = max_iter; j->rc = 0;
    if (t > 0) pthread_create(&th[t], NULL, worker, j);
  }
  worker(&jobs[0]);
  int rc = jobs[0].rc;
  bf_counts total = jobs[0].cnt;
  for (int t = 1; t < threads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
    total.ready += jobs[t].cnt.ready; total.skip += jobs[t].cnt.skip;
    total.expansion += jobs[t].cnt.expansion; total.evals += jobs[t].cnt.evals;
  }
  if (counts) *counts = total;
  free(jobs); free(th);
  return rc;
}

/* Expansion tuples for a finished batch, in (run, step, branch) order — the order
 * executeParallelStep creates children (step_executor.go:750-806). */
int orc_packed_expand(const orc_topology* topos, uint32_t n_topos, const bf_layout* L, uint32_t n_runs,
                      const uint8_t* state, const uint8_t* result, bf_expansion* out, uint64_t cap,
                      uint64_t* n_out) {
  uint64_t n = 0;
  for (uint32_t r = 0; r < n_runs; ++r) {
    const bf_run_header* rh = (const bf_run_header*)(state + (size_t)r * L->state_stride);
    if (rh->topo_slot >= n_topos) return BF_ETOPO;
    const orc_topology* T = &topos[rh->topo_slot];
    const uint32_t* ready = (const uint32_t*)(result + (size_t)r * L->result_stride + L->off_ready);
    for (uint32_t q = 0; q < T->n_parallel; ++q) {
      if (!tst_mask(ready, T->parallel[q].step)) continue;
      for (uint32_t b = 0; b < T->parallel[q].branches; ++b) {
        if (out && n < cap) { out[n].run = r; out[n].step = T->parallel[q].step; out[n].branch = (uint16_t)b; }
        n++;
      }
    }
  }
  if (n_out) *n_out = n;
  return BF_OK;
}
