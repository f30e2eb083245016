// host_mirror.cc — C++ host-side mirror of the reference's Story / StoryRun objects for the frontier path
// (include/bobrafrontier_host.h).  Pure host code: name -> index maps, the template-reference scanner of
// buildDependencyGraphs, CSR packing, in-place bit-plane updates of the state records, result decoding.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bobrafrontier_host.h"

namespace {

inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
inline bool name_char(char c) {  // [a-zA-Z0-9_\-], the character class of dag.go:3029
  return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-';
}
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// One left-to-right pass equivalent to regexp.FindAllStringSubmatch with the pattern of dag.go:3028-3030:
//   steps\.(NAME)\.  |  steps\s*\[\s*['"](NAME)['"]\s*\]  |  \(index\s+\.steps\s+["'](NAME)["']\)
// (leftmost match, alternatives tried in order, matches do not overlap).
void scan_step_refs(const std::string& e, std::vector<std::string>& out) {
  const size_t n = e.size();
  size_t i = 0;
  auto name_at = [&](size_t p, size_t& end) {  // greedy NAME+ starting at p
    size_t q = p;
    while (q < n && name_char(e[q])) ++q;
    end = q;
    return q > p;
  };
  while (i < n) {
    size_t adv = 0;
    if (e.compare(i, 5, "steps") == 0) {
      size_t p = i + 5, q;
      // alternative 1: steps.NAME.
      if (p < n && e[p] == '.' && name_at(p + 1, q) && q < n && e[q] == '.') {
        out.emplace_back(e, p + 1, q - p - 1);
        adv = q + 1 - i;
      } else {
        // alternative 2: steps\s*[\s*'NAME'\s*]
        size_t r = p;
        while (r < n && is_space(e[r])) ++r;
        if (r < n && e[r] == '[') {
          ++r;
          while (r < n && is_space(e[r])) ++r;
          if (r < n && (e[r] == '\'' || e[r] == '"')) {
            size_t s0 = r + 1, s1;
            if (name_at(s0, s1) && s1 < n && (e[s1] == '\'' || e[s1] == '"')) {
              size_t t = s1 + 1;
              while (t < n && is_space(e[t])) ++t;
              if (t < n && e[t] == ']') {
                out.emplace_back(e, s0, s1 - s0);
                adv = t + 1 - i;
              }
            }
          }
        }
      }
    } else if (e.compare(i, 6, "(index") == 0) {
      // alternative 3: (index\s+.steps\s+"NAME")
      size_t r = i + 6, r0 = r;
      while (r < n && is_space(e[r])) ++r;
      if (r > r0 && e.compare(r, 6, ".steps") == 0) {
        size_t u = r + 6, u0 = u;
        while (u < n && is_space(e[u])) ++u;
        if (u > u0 && u < n && (e[u] == '"' || e[u] == '\'')) {
          size_t s0 = u + 1, s1;
          if (name_at(s0, s1) && s1 + 1 < n && (e[s1] == '"' || e[s1] == '\'') && e[s1 + 1] == ')') {
            out.emplace_back(e, s0, s1 - s0);
            adv = s1 + 2 - i;
          }
        }
      }
    }
    i += adv ? adv : 1;
  }
}

std::string sanitize(const std::string& name) {  // step_executor.go:1652-1670
  std::string s = name;
  for (char& c : s)
    if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_')) c = '_';
  return s;
}

const char* kPhaseNames[14] = {"",       "Pending", "Running",    "Succeeded", "Failed",  "Finished", "Canceled",
                               "Compensated", "Paused",  "Blocked", "Scheduling", "Timeout", "Aborted", "Skipped"};
const char* kQueuedPrefixes[4] = {"Queued due to story concurrency limit", "Queued due to queue concurrency limit",
                                  "Queued due to global concurrency limit", "Queued due to higher-priority work"};

int phase_code_of(const char* phase, const char* message) {
  if (!phase || !*phase) return BF_PHASE_NONE;
  for (int c = 1; c <= 13; ++c)
    if (!strcmp(phase, kPhaseNames[c])) {
      if (c == BF_PHASE_PENDING && message)
        for (const char* p : kQueuedPrefixes)
          if (!strncmp(message, p, strlen(p))) return BF_PHASE_PENDING_QUEUED;  // dag.go:2035-2051
      return c;
    }
  return -1;
}

struct StepDef {
  std::string name;
  int group, type;
  bool allow, toskip, has_if, has_with;
  std::string if_expr, with_raw;
  std::vector<std::string> needs;
  std::vector<std::pair<std::string, bool>> branches;
};

inline void set_code(uint8_t* base, uint32_t W, int nbits, uint32_t i, int v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(base);
  const uint32_t m = 1u << (i & 31u);
  for (int b = 0; b < nbits; ++b) {
    uint32_t& x = w[(uint32_t)b * W + (i >> 5)];
    x = ((v >> b) & 1) ? (x | m) : (x & ~m);
  }
}
inline int get_code(const uint8_t* base, uint32_t W, int nbits, uint32_t i) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base);
  int v = 0;
  for (int b = 0; b < nbits; ++b) v |= (int)((w[(uint32_t)b * W + (i >> 5)] >> (i & 31u)) & 1u) << b;
  return v;
}

}  // namespace

struct bfh_story {
  std::vector<StepDef> defs;           // insertion order
  int continue_on_failure = -1;
  bool realtime = false;
  bool finalized = false;
  std::string err;
  // packed form (index space = allStorySteps)
  std::vector<uint32_t> order;         // index -> def handle
  std::vector<uint32_t> index_of_def;  // def handle -> index
  std::unordered_map<std::string, uint32_t> index;
  std::vector<uint32_t> row_ptr;
  std::vector<uint16_t> col_idx;
  std::vector<uint8_t> flags;
  std::vector<bf_parallel_desc> par;
  std::vector<uint8_t> allow_bits;
  std::vector<uint32_t> child_first;
  uint32_t child_nibbles = 0;
};

struct bfh_batch {
  bf_ctx* ctx = nullptr;
  bf_layout L{};
  uint32_t cap = 0, n = 0;
  uint8_t* state = nullptr;
  uint8_t* result = nullptr;
  bool pinned = false;
  std::vector<const bfh_story*> story_of_run;
  std::string err;
  // resident mode (row f2): the device holds the records, setters log coalesced deltas, eval sends only those
  bool resident = false;
  uint32_t handle = 0;
  uint32_t uploaded = 0;                             // runs [0, uploaded) exist on the device
  std::unordered_map<uint64_t, uint32_t> delta_at;   // (run, field, index) -> position in `deltas`
  std::vector<bf_delta> deltas;
  uint64_t delta_bytes_sent = 0, full_bytes_sent = 0;

  void log_delta(uint32_t run, uint32_t field, uint32_t index, uint32_t code) {
    if (!resident || run >= uploaded) return;        // a run not yet on the device travels as a full record
    const uint64_t key = ((uint64_t)run << 32) | ((uint64_t)field << 16) | index;
    auto it = delta_at.find(key);
    if (it != delta_at.end()) { deltas[it->second].code = (uint8_t)code; return; }   // the last value of a tick wins
    delta_at[key] = (uint32_t)deltas.size();
    bf_delta d;
    d.run = run; d.index = (uint16_t)index; d.field = (uint8_t)field; d.code = (uint8_t)code;
    deltas.push_back(d);
  }
};

extern "C" {

// ------------------------------------------------------------------------------------------ Story
bfh_story* bfh_story_new(void) { return new (std::nothrow) bfh_story(); }
void bfh_story_free(bfh_story* s) { delete s; }
const char* bfh_story_error(const bfh_story* s) { return s ? s->err.c_str() : "null story"; }

int bfh_story_add_step(bfh_story* s, const char* name, int group, int type, int allow_failure, int on_timeout_skip,
                       const char* if_expr, const char* with_raw) {
  if (!s || !name || group < 0 || group > 2 || type < 0 || type > 7) return BF_EINVAL;
  if (s->finalized) { s->err = "story already finalized"; return BF_EINVAL; }
  StepDef d;
  d.name = name; d.group = group; d.type = type; d.allow = allow_failure != 0; d.toskip = on_timeout_skip != 0;
  d.has_if = if_expr != nullptr; d.has_with = with_raw != nullptr;
  if (if_expr) d.if_expr = if_expr;
  if (with_raw) d.with_raw = with_raw;
  s->defs.push_back(std::move(d));
  return (int)s->defs.size() - 1;
}
int bfh_step_add_need(bfh_story* s, int step, const char* dep) {
  if (!s || !dep || step < 0 || (size_t)step >= s->defs.size() || s->finalized) return BF_EINVAL;
  s->defs[step].needs.emplace_back(dep);
  return BF_OK;
}
int bfh_step_add_branch(bfh_story* s, int step, const char* bname, int allow) {
  if (!s || !bname || step < 0 || (size_t)step >= s->defs.size() || s->finalized) return BF_EINVAL;
  s->defs[step].branches.emplace_back(bname, allow != 0);
  return BF_OK;
}
int bfh_story_set_policy(bfh_story* s, int cont, int realtime) {
  if (!s) return BF_EINVAL;
  s->continue_on_failure = cont;
  s->realtime = realtime != 0;
  return BF_OK;
}

int bfh_story_finalize(bfh_story* s) {
  if (!s) return BF_EINVAL;
  if (s->finalized) return BF_OK;
  const uint32_t n = (uint32_t)s->defs.size();
  if (n == 0 || n > BF_MAX_STEPS) { s->err = "step count out of range (1..1024)"; return BF_ETOPO; }
  // allStorySteps order: main ++ compensations ++ finally (dag.go:3270-3280)
  s->order.clear();
  for (int g = 0; g < 3; ++g)
    for (uint32_t h = 0; h < n; ++h)
      if (s->defs[h].group == g) s->order.push_back(h);
  s->index_of_def.assign(n, 0);
  s->index.clear();
  for (uint32_t i = 0; i < n; ++i) {
    s->index_of_def[s->order[i]] = i;
    if (!s->index.emplace(s->defs[s->order[i]].name, i).second) { s->err = "duplicate step name " + s->defs[s->order[i]].name; return BF_ETOPO; }
  }
  // buildDependencyGraphs is called on the evaluated group's list (dag.go:1700): alias maps are per group
  std::vector<std::vector<uint32_t>> rows(n);
  std::vector<std::string> unknown;
  for (int g = 0; g < 3; ++g) {
    std::unordered_map<std::string, std::string> alias_to_real;  // dag.go:3033-3039
    for (uint32_t i = 0; i < n; ++i) {
      const StepDef& d = s->defs[s->order[i]];
      if (d.group != g) continue;
      const std::string a = sanitize(d.name);
      if (a != d.name) alias_to_real[a] = d.name;
    }
    for (uint32_t i = 0; i < n; ++i) {
      const StepDef& d = s->defs[s->order[i]];
      if (d.group != g) continue;
      std::vector<std::string> deps = d.needs;  // 1. explicit needs (dag.go:3051)
      std::vector<std::string> refs;
      if (d.has_if) scan_step_refs(d.if_expr, refs);  // 2. `if` (dag.go:3056)
      if (d.has_with && (d.type == BF_STEP_ENGRAM || d.type == BF_STEP_EXECUTE_STORY)) scan_step_refs(d.with_raw, refs);  // 3. :3061-3070
      for (std::string& r : refs) {
        auto it = alias_to_real.find(r);
        deps.push_back(it == alias_to_real.end() ? r : it->second);
      }
      for (const std::string& dep : deps) {
        auto it = s->index.find(dep);
        if (it == s->index.end()) unknown.push_back(d.name + "->" + dep);
        else rows[i].push_back(it->second);
      }
      std::sort(rows[i].begin(), rows[i].end());
      rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
    }
  }
  if (!unknown.empty()) {  // dag.go:3087-3098
    std::sort(unknown.begin(), unknown.end());
    s->err = "unknown step dependencies: ";
    for (size_t k = 0; k < unknown.size(); ++k) s->err += (k ? ", " : "") + unknown[k];
    return BF_ETOPO;
  }
  s->row_ptr.assign(n + 1, 0);
  s->col_idx.clear();
  s->flags.assign(n, 0);
  s->par.clear(); s->allow_bits.clear(); s->child_first.clear();
  std::vector<bool> bits;
  uint32_t nib = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const StepDef& d = s->defs[s->order[i]];
    s->row_ptr[i + 1] = s->row_ptr[i] + (uint32_t)rows[i].size();
    for (uint32_t c : rows[i]) s->col_idx.push_back((uint16_t)c);
    uint8_t f = (uint8_t)d.type;
    if (d.allow) f |= BF_SF_ALLOW_FAILURE;
    if (d.toskip && (d.type == BF_STEP_GATE || d.type == BF_STEP_WAIT)) f |= BF_SF_ON_TIMEOUT_SKIP;
    if (d.has_if && !d.if_expr.empty()) f |= BF_SF_HAS_IF;
    f |= (uint8_t)(d.group << BF_SF_GROUP_SHIFT);
    s->flags[i] = f;
    if (d.type == BF_STEP_PARALLEL) {
      if (s->par.size() >= BF_MAX_PARALLEL) { s->err = "more than 64 parallel steps"; return BF_ETOPO; }
      bf_parallel_desc pd;
      pd.step = (uint16_t)i; pd.branches = (uint16_t)d.branches.size(); pd.allow_first = (uint32_t)bits.size();
      for (auto& b : d.branches) bits.push_back(b.second);
      s->par.push_back(pd);
      nib = round_up(nib, 8);
      s->child_first.push_back(nib);
      nib += pd.branches;
    }
  }
  if (s->col_idx.size() > BF_MAX_EDGES) { s->err = "too many edges"; return BF_ETOPO; }
  s->child_nibbles = round_up(nib, 8);
  s->allow_bits.assign((bits.size() + 7) / 8, 0);
  for (size_t k = 0; k < bits.size(); ++k)
    if (bits[k]) s->allow_bits[k >> 3] |= (uint8_t)(1u << (k & 7));
  s->finalized = true;
  return BF_OK;
}

int bfh_story_dims(const bfh_story* s, uint32_t* S, uint32_t* E, uint32_t* P) {
  if (!s || !s->finalized) return BF_EINVAL;
  if (S) *S = (uint32_t)s->flags.size();
  if (E) *E = (uint32_t)s->col_idx.size();
  if (P) *P = (uint32_t)s->par.size();
  return BF_OK;
}
int bfh_story_csr(const bfh_story* s, uint32_t* rp, uint16_t* ci, uint8_t* fl) {
  if (!s || !s->finalized) return BF_EINVAL;
  if (rp) memcpy(rp, s->row_ptr.data(), s->row_ptr.size() * 4);
  if (ci && !s->col_idx.empty()) memcpy(ci, s->col_idx.data(), s->col_idx.size() * 2);
  if (fl) memcpy(fl, s->flags.data(), s->flags.size());
  return BF_OK;
}
int bfh_story_upload(bfh_story* s, bf_ctx* ctx, uint32_t* slot_out) {
  if (!s || !ctx || !slot_out) return BF_EINVAL;
  if (!s->finalized)
    if (int rc = bfh_story_finalize(s)) return rc;
  bf_topology t{};
  t.n_steps = (uint32_t)s->flags.size(); t.n_edges = (uint32_t)s->col_idx.size();
  t.row_ptr = s->row_ptr.data(); t.col_idx = s->col_idx.data(); t.step_flags = s->flags.data();
  t.parallel = s->par.data(); t.n_parallel = (uint32_t)s->par.size();
  t.branch_allow_bits = s->allow_bits.empty() ? nullptr : s->allow_bits.data();
  t.n_branch_allow_bits = (uint32_t)s->allow_bits.size() * 8;
  const int rc = bf_topology_put(ctx, &t, slot_out);
  if (rc != BF_OK) s->err = bf_last_error(ctx);
  return rc;
}
int bfh_story_step_index(const bfh_story* s, const char* name) {
  if (!s || !s->finalized || !name) return -1;
  auto it = s->index.find(name);
  return it == s->index.end() ? -1 : (int)it->second;
}
const char* bfh_story_step_name(const bfh_story* s, uint32_t idx) {
  if (!s || !s->finalized || idx >= s->order.size()) return nullptr;
  return s->defs[s->order[idx]].name.c_str();
}
uint32_t bfh_story_run_flags(const bfh_story* s) {
  if (!s) return 0;
  uint32_t f = 0;
  if (s->continue_on_failure != 1) f |= BF_RF_FAIL_FAST;  // shouldFailFast, dag.go:3504-3511
  if (s->realtime) f |= BF_RF_REALTIME;
  return f;
}
int bfh_scan_step_refs(const char* expression, char* out, size_t cap) {
  if (!expression) return BF_EINVAL;
  std::vector<std::string> refs;
  scan_step_refs(expression, refs);
  std::string j;
  for (auto& r : refs) { j += r; j += '\n'; }
  if (out && cap) { strncpy(out, j.c_str(), cap - 1); out[cap - 1] = 0; }
  return (int)refs.size();
}

// ------------------------------------------------------------------------------------------ Batch
bfh_batch* bfh_batch_new(bf_ctx* ctx, uint32_t steps_max, uint32_t child_nibbles, uint32_t fields, uint32_t capacity) {
  bfh_batch* b = new (std::nothrow) bfh_batch();
  if (!b) return nullptr;
  b->ctx = ctx;
  if (bf_layout_init(&b->L, steps_max, child_nibbles, fields) != BF_OK || capacity == 0) { delete b; return nullptr; }
  b->cap = capacity;
  const size_t sb = (size_t)capacity * b->L.state_stride, rb = (size_t)capacity * b->L.result_stride;
  void *ps = nullptr, *pr = nullptr;
  if (ctx && bf_alloc_pinned(ctx, sb, &ps) == BF_OK && bf_alloc_pinned(ctx, rb, &pr) == BF_OK) {
    b->pinned = true;
  } else {
    if (ctx && ps) bf_free_pinned(ctx, ps);
    ps = malloc(sb); pr = malloc(rb);
    if (!ps || !pr) { free(ps); free(pr); delete b; return nullptr; }
  }
  b->state = static_cast<uint8_t*>(ps); b->result = static_cast<uint8_t*>(pr);
  memset(b->state, 0, sb); memset(b->result, 0, rb);
  return b;
}
void bfh_batch_free(bfh_batch* b) {
  if (!b) return;
  if (b->resident) bf_resident_destroy(b->ctx, b->handle);
  if (b->pinned) { bf_free_pinned(b->ctx, b->state); bf_free_pinned(b->ctx, b->result); }
  else { free(b->state); free(b->result); }
  delete b;
}
const char* bfh_batch_error(const bfh_batch* b) { return b ? b->err.c_str() : "null batch"; }
const bf_layout* bfh_batch_layout(const bfh_batch* b) { return b ? &b->L : nullptr; }
uint32_t bfh_batch_size(const bfh_batch* b) { return b ? b->n : 0; }
const void* bfh_batch_state(const bfh_batch* b) { return b ? b->state : nullptr; }
const void* bfh_batch_result(const bfh_batch* b) { return b ? b->result : nullptr; }

int bfh_batch_add_run(bfh_batch* b, const bfh_story* story, uint32_t slot) {
  if (!b || !story || !story->finalized) return BF_EINVAL;
  if (b->n >= b->cap) { b->err = "batch full"; return BF_ENOMEM; }
  if (story->flags.size() > b->L.steps_max) { b->err = "story larger than the batch layout"; return BF_EINVAL; }
  if (b->L.off_child != BF_OFF_NONE && story->child_nibbles > b->L.child_nibbles) { b->err = "child area too small"; return BF_EINVAL; }
  uint8_t* rec = b->state + (size_t)b->n * b->L.state_stride;
  memset(rec, 0, b->L.state_stride);
  bf_run_header* h = reinterpret_cast<bf_run_header*>(rec);
  h->topo_slot = slot;
  h->run_flags = (uint8_t)bfh_story_run_flags(story);
  b->story_of_run.push_back(story);
  return (int)b->n++;
}
int bfh_batch_remove_last_run(bfh_batch* b) {
  if (!b || b->n == 0) return BF_EINVAL;
  b->story_of_run.pop_back();
  --b->n;
  if (b->uploaded > b->n) b->uploaded = b->n;  // the slot's next occupant is uploaded as a full record
  // pending deltas of the removed run must not land on the slot's next occupant (whose full record is uploaded first)
  if (b->resident && !b->deltas.empty()) {
    const uint32_t gone = b->n;
    size_t keep = 0;
    for (size_t i = 0; i < b->deltas.size(); ++i)
      if (b->deltas[i].run != gone) b->deltas[keep++] = b->deltas[i];
    if (keep != b->deltas.size()) {
      b->deltas.resize(keep);
      b->delta_at.clear();
      for (size_t i = 0; i < keep; ++i) {
        const bf_delta& d = b->deltas[i];
        b->delta_at[((uint64_t)d.run << 32) | ((uint64_t)d.field << 16) | d.index] = (uint32_t)i;
      }
    }
  }
  return BF_OK;
}

#define RUN_CHECK(b, run, step)                                                                                   \
  if (!(b) || (run) >= (b)->n) return BF_EINVAL;                                                                 \
  const bfh_story* st = (b)->story_of_run[run];                                                                  \
  if ((step) >= st->flags.size()) return BF_EINVAL;                                                              \
  uint8_t* rec = (b)->state + (size_t)(run) * (b)->L.state_stride;

int bfh_run_set_phase_code(bfh_batch* b, uint32_t run, uint32_t step, int code) {
  RUN_CHECK(b, run, step)
  if (code < 0 || code > 14) return BF_EINVAL;
  set_code(rec + b->L.off_phase, b->L.words, 4, step, code);
  b->log_delta(run, BF_DELTA_PHASE, step, (uint32_t)code);
  return BF_OK;
}
int bfh_run_set_phase(bfh_batch* b, uint32_t run, uint32_t step, const char* phase, const char* message) {
  const int c = phase_code_of(phase, message);
  if (c < 0) return BF_EINVAL;
  return bfh_run_set_phase_code(b, run, step, c);
}
int bfh_run_set_cond(bfh_batch* b, uint32_t run, uint32_t step, int code) {
  RUN_CHECK(b, run, step)
  if (b->L.off_cond == BF_OFF_NONE || code < 0 || code > 3) return BF_EINVAL;
  set_code(rec + b->L.off_cond, b->L.words, 2, step, code);
  b->log_delta(run, BF_DELTA_COND, step, (uint32_t)code);
  return BF_OK;
}
int bfh_run_set_decision(bfh_batch* b, uint32_t run, uint32_t step, int code) {
  RUN_CHECK(b, run, step)
  if (b->L.off_decision == BF_OFF_NONE || code < 0 || code > 3) return BF_EINVAL;
  set_code(rec + b->L.off_decision, b->L.words, 2, step, code);
  b->log_delta(run, BF_DELTA_DECISION, step, (uint32_t)code);
  return BF_OK;
}
int bfh_run_set_gate(bfh_batch* b, uint32_t run, uint32_t step, const char* gate_state, int timed_out) {
  int d = BF_DEC_PENDING;  // dag.go:1489-1533: Approved > Rejected > (pending: timeout check)
  if (gate_state && !strcmp(gate_state, "Approved")) d = BF_DEC_SUCCEED;
  else if (gate_state && !strcmp(gate_state, "Rejected")) d = BF_DEC_FAIL;
  else if (timed_out) d = BF_DEC_TIMED_OUT;
  return bfh_run_set_decision(b, run, step, d);
}
int bfh_run_set_run_flags(bfh_batch* b, uint32_t run, int topology_terminated, int host_group) {
  if (!b || run >= b->n) return BF_EINVAL;
  bf_run_header* h = reinterpret_cast<bf_run_header*>(b->state + (size_t)run * b->L.state_stride);
  uint8_t f = (uint8_t)bfh_story_run_flags(b->story_of_run[run]);
  if (topology_terminated) f |= BF_RF_TOPOLOGY_TERMINATED;
  if (host_group >= 0) f |= (uint8_t)(BF_RF_HOST_GROUP | ((host_group & 3) << BF_RF_HOST_GROUP_SHIFT));
  h->run_flags = f;
  b->log_delta(run, BF_DELTA_RUN_FLAGS, 0, f);
  return BF_OK;
}
int bfh_run_register_children(bfh_batch* b, uint32_t run, uint32_t q, int on) {
  if (!b || run >= b->n || q >= b->story_of_run[run]->par.size()) return BF_EINVAL;
  bf_run_header* h = reinterpret_cast<bf_run_header*>(b->state + (size_t)run * b->L.state_stride);
  if (on) h->children_registered |= 1ull << q; else h->children_registered &= ~(1ull << q);
  b->log_delta(run, BF_DELTA_REGISTERED, q, on ? 1u : 0u);
  return BF_OK;
}
int bfh_run_set_child_phase(bfh_batch* b, uint32_t run, uint32_t q, uint32_t branch, const char* phase) {
  if (!b || run >= b->n || b->L.off_child == BF_OFF_NONE) return BF_EINVAL;
  const bfh_story* st = b->story_of_run[run];
  if (q >= st->par.size() || branch >= st->par[q].branches) return BF_EINVAL;
  const int c = phase_code_of(phase, nullptr);
  if (c < 0) return BF_EINVAL;
  uint8_t* ch = b->state + (size_t)run * b->L.state_stride + b->L.off_child;
  const uint32_t i = st->child_first[q] + branch;
  const uint8_t sh = (uint8_t)((i & 1u) * 4u);
  ch[i >> 1] = (uint8_t)((ch[i >> 1] & ~(0xFu << sh)) | ((unsigned)c << sh));
  b->log_delta(run, BF_DELTA_CHILD, i, (uint32_t)c);
  return BF_OK;
}

int bfh_batch_eval(bfh_batch* b, uint32_t eval_flags, bf_counts* counts) {
  if (!b) return BF_EINVAL;
  if (!b->ctx) { b->err = "batch was created without a device context"; return BF_ENODEV; }
  if (b->resident) {
    // new runs as full records, everything else as the tick's coalesced deltas, then one pass over the device copy
    int rc = BF_OK;
    if (b->uploaded < b->n) {
      rc = bf_resident_upload(b->ctx, b->handle, b->uploaded, b->n - b->uploaded, b->state + (size_t)b->uploaded * b->L.state_stride);
      if (rc == BF_OK) { b->full_bytes_sent += (uint64_t)(b->n - b->uploaded) * b->L.state_stride; b->uploaded = b->n; }
    }
    if (rc == BF_OK && !b->deltas.empty()) {
      rc = bf_resident_apply(b->ctx, b->handle, b->deltas.data(), (uint32_t)b->deltas.size());
      b->delta_bytes_sent += b->deltas.size() * sizeof(bf_delta);
    }
    b->deltas.clear();
    b->delta_at.clear();
    if (rc == BF_OK) rc = bf_resident_eval(b->ctx, b->handle, b->n, eval_flags & ~(uint32_t)BF_EVAL_EXPANSION, 0, b->result, counts);
    if (rc != BF_OK) b->err = bf_last_error(b->ctx);
    return rc;
  }
  bf_batch bb{};
  bb.struct_size = sizeof(bf_batch);
  bb.n_runs = b->n; bb.flags = eval_flags; bb.layout = b->L;
  bb.state = b->state; bb.result = b->result; bb.counts = counts;
  const int rc = bf_eval(b->ctx, &bb);
  if (rc != BF_OK) b->err = bf_last_error(b->ctx);
  return rc;
}

int bfh_batch_set_resident(bfh_batch* b, int on) {
  if (!b) return BF_EINVAL;
  if (!b->ctx) { b->err = "batch was created without a device context"; return BF_ENODEV; }
  if ((on != 0) == b->resident) return BF_OK;
  if (on) {
    const int rc = bf_resident_create(b->ctx, &b->L, b->cap, &b->handle);
    if (rc != BF_OK) { b->err = bf_last_error(b->ctx); return rc; }
    b->resident = true; b->uploaded = 0;
  } else {
    bf_resident_destroy(b->ctx, b->handle);
    b->resident = false; b->uploaded = 0;
    b->deltas.clear(); b->delta_at.clear();
  }
  return BF_OK;
}
int bfh_batch_traffic(const bfh_batch* b, uint64_t* full_record_bytes, uint64_t* delta_bytes, uint32_t* pending_deltas) {
  if (!b) return BF_EINVAL;
  if (full_record_bytes) *full_record_bytes = b->full_bytes_sent;
  if (delta_bytes) *delta_bytes = b->delta_bytes_sent;
  if (pending_deltas) *pending_deltas = (uint32_t)b->deltas.size();
  return BF_OK;
}

uint32_t bfh_run_summary(const bfh_batch* b, uint32_t run) {
  if (!b || run >= b->n) return 0xFFFFFFFFu;
  return reinterpret_cast<const bf_result_header*>(b->result + (size_t)run * b->L.result_stride)->summary;
}
static int mask_list(const bfh_batch* b, uint32_t run, uint32_t off, uint32_t* out, uint32_t cap) {
  if (!b || run >= b->n || off == BF_OFF_NONE) return BF_EINVAL;
  const uint32_t* m = reinterpret_cast<const uint32_t*>(b->result + (size_t)run * b->L.result_stride + off);
  const uint32_t S = (uint32_t)b->story_of_run[run]->flags.size();
  int n = 0;
  for (uint32_t i = 0; i < S; ++i)
    if ((m[i >> 5] >> (i & 31u)) & 1u) {
      if (out && (uint32_t)n < cap) out[n] = i;
      ++n;
    }
  return n;
}
int bfh_run_ready(const bfh_batch* b, uint32_t run, uint32_t* out, uint32_t cap) { return mask_list(b, run, b ? b->L.off_ready : 0, out, cap); }
int bfh_run_skipped(const bfh_batch* b, uint32_t run, uint32_t* out, uint32_t cap) { return mask_list(b, run, b ? b->L.off_skip : 0, out, cap); }
int bfh_run_failed(const bfh_batch* b, uint32_t run, uint32_t* out, uint32_t cap) { return mask_list(b, run, b ? b->L.off_fail : BF_OFF_NONE, out, cap); }
int bfh_run_needs_cond(const bfh_batch* b, uint32_t run, uint32_t* out, uint32_t cap) { return mask_list(b, run, b ? b->L.off_needs_cond : BF_OFF_NONE, out, cap); }
int bfh_run_phase_out(const bfh_batch* b, uint32_t run, uint32_t step) {
  if (!b || run >= b->n || b->L.off_phase_out == BF_OFF_NONE || step >= b->story_of_run[run]->flags.size()) return -1;
  return get_code(b->result + (size_t)run * b->L.result_stride + b->L.off_phase_out, b->L.words, 4, step);
}
int bfh_run_skip_reason(const bfh_batch* b, uint32_t run, uint32_t step, char* out, size_t cap) {
  if (!b || run >= b->n || !out || !cap) return BF_EINVAL;
  const bfh_story* st = b->story_of_run[run];
  if (step >= st->flags.size()) return BF_EINVAL;
  const uint8_t* rrec = b->result + (size_t)run * b->L.result_stride;
  std::string msg = "Skipped due to 'if' condition";  // dag.go:2831
  bool by_dep = false;
  if (b->L.off_skip_dep != BF_OFF_NONE)
    by_dep = (reinterpret_cast<const uint32_t*>(rrec + b->L.off_skip_dep)[step >> 5] >> (step & 31u)) & 1u;
  if (by_dep) {
    msg = "Skipped due to failed dependency: ";  // dag.go:2737; the named dep = first failed one in CSR order
    const uint8_t* ph = b->L.off_phase_out != BF_OFF_NONE ? rrec + b->L.off_phase_out
                                                         : b->state + (size_t)run * b->L.state_stride + b->L.off_phase;
    for (uint32_t e = st->row_ptr[step]; e < st->row_ptr[step + 1]; ++e) {
      const uint32_t d = st->col_idx[e];
      const int p = get_code(ph, b->L.words, 4, d);
      const bool term = (BF_LUT_TERMINAL >> p) & 1u;
      const bool compl_ = ((BF_LUT_COMPLETED0 >> p) & 1u) || (term && (st->flags[d] & BF_SF_ALLOW_FAILURE));
      // a dep set Failed earlier in this same pass (dag.go:2745) counts even with allowFailure: `completed` is stale
      const bool failed_now = b->L.off_fail != BF_OFF_NONE && d < step &&
                              ((reinterpret_cast<const uint32_t*>(rrec + b->L.off_fail)[d >> 5] >> (d & 31u)) & 1u);
      if ((term && !compl_) || failed_now) { msg += st->defs[st->order[d]].name; break; }
    }
  }
  strncpy(out, msg.c_str(), cap - 1);
  out[cap - 1] = 0;
  return BF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ limiters (rows a9 / f4)
namespace {
std::string normalize_queue(const char* name) {  // scheduling.go:21-27: TrimSpace + ToLower
  std::string q = name ? name : "";
  size_t a = 0, b = q.size();
  while (a < b && is_space(q[a])) ++a;
  while (b > a && is_space(q[b - 1])) --b;
  q = q.substr(a, b - a);
  for (char& c : q) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
  return q;
}
int phase_code_plain(const char* phase) { return phase_code_of(phase ? phase : "", ""); }
}  // namespace

struct bfh_sched {
  bfh_batch* batch = nullptr;
  std::string err;
  int32_t global_limit = 0;
  uint32_t global_base = 0;
  std::vector<std::string> queue_names;
  std::unordered_map<std::string, uint32_t> queue_key;
  std::vector<int32_t> queue_limit, queue_default_prio, queue_aging;
  std::vector<uint32_t> queue_base;
  std::unordered_map<std::string, uint32_t> story_key;   // "namespace/name"
  std::vector<int32_t> story_limit;
  std::vector<uint32_t> story_base;
  std::vector<bf_sched_run> runs;
  // results
  std::vector<uint8_t> records;
  std::vector<uint32_t> story_running, queue_running;
  std::vector<int32_t> queue_maxprio;
  uint32_t global_running = 0;
  bool applied = false;

  uint32_t queue_of(const std::string& normalized) {  // "" -> default (queueLabelValue, scheduling.go:29-35)
    const std::string n = normalized.empty() ? "default" : normalized;
    auto it = queue_key.find(n);
    if (it != queue_key.end()) return it->second;
    const uint32_t k = (uint32_t)queue_names.size();
    queue_key[n] = k;
    queue_names.push_back(n);
    queue_limit.push_back(0); queue_default_prio.push_back(0); queue_aging.push_back(0); queue_base.push_back(0);
    return k;
  }
  uint32_t story_of(const char* ns, const char* name) {
    const std::string k = std::string(ns ? ns : "") + "/" + (name ? name : "");
    auto it = story_key.find(k);
    if (it != story_key.end()) return it->second;
    const uint32_t id = (uint32_t)story_limit.size();
    story_key[k] = id;
    story_limit.push_back(0); story_base.push_back(0);
    return id;
  }
};

extern "C" {

bfh_sched* bfh_sched_new(bfh_batch* batch) {
  if (!batch) return nullptr;
  bfh_sched* s = new (std::nothrow) bfh_sched();
  if (!s) return nullptr;
  s->batch = batch;
  // the default scheduling config: queue "default" ages every 60 s (internal/config/controller_config.go:731-739)
  const uint32_t k = s->queue_of("default");
  s->queue_aging[k] = 60;
  return s;
}
void bfh_sched_free(bfh_sched* s) { delete s; }
const char* bfh_sched_error(const bfh_sched* s) { return s ? s->err.c_str() : "null sched"; }

int bfh_sched_set_global(bfh_sched* s, int32_t limit, uint32_t base) {
  if (!s) return BF_EINVAL;
  s->global_limit = limit; s->global_base = base;
  return BF_OK;
}
int bfh_sched_set_queue(bfh_sched* s, const char* name, int32_t concurrency, int32_t default_priority, int32_t aging, uint32_t base) {
  if (!s) return BF_EINVAL;
  const uint32_t k = s->queue_of(normalize_queue(name));
  s->queue_limit[k] = concurrency; s->queue_default_prio[k] = default_priority; s->queue_aging[k] = aging; s->queue_base[k] = base;
  return (int)k;
}
int bfh_sched_set_story_base(bfh_sched* s, const char* ns, const char* name, uint32_t base) {
  if (!s) return BF_EINVAL;
  const uint32_t k = s->story_of(ns, name);
  s->story_base[k] = base;
  return (int)k;
}
int bfh_sched_set_run(bfh_sched* s, uint32_t run, const char* ns, const char* name, int32_t story_concurrency, const char* policy_queue,
                      int has_priority, int32_t policy_priority, const char* run_phase, int64_t queued_since_ns, int64_t now_ns) {
  if (!s || run >= s->batch->n) { if (s) s->err = "run index out of range"; return BF_EINVAL; }
  if (s->runs.size() < s->batch->n) s->runs.resize(s->batch->n);
  // resolveSchedulingDecision, scheduling.go:130-163 (no fallback decision on this path: dag.go:1805 passes nil)
  const std::string sq = normalize_queue(policy_queue);
  const uint32_t qk = s->queue_of(sq);
  int32_t prio;
  if (has_priority) prio = policy_priority;
  else prio = s->queue_default_prio[qk];  // both remaining cases read queueDefaultPriority of the decided queue
  const uint32_t sk = s->story_of(ns, name);
  s->story_limit[sk] = story_concurrency > 0 ? story_concurrency : 0;  // storyConcurrencyLimit, dag.go:1863-1868
  bf_sched_run& r = s->runs[run];
  memset(&r, 0, sizeof r);
  r.story_key = sk; r.queue_key = qk; r.priority = prio;
  r.run_phase = (uint32_t)phase_code_plain(run_phase);
  if (queued_since_ns < 0) r.queued_elapsed_s = BF_SCHED_NONE;
  else {
    const int64_t el_ns = now_ns - queued_since_ns;  // elapsed <= 0 -> base priority (dag.go:1952-1955)
    const int64_t el = el_ns <= 0 ? 0 : el_ns / 1000000000ll;  // int32(elapsed.Seconds())
    r.queued_elapsed_s = el > 0x7FFFFFFF ? 0x7FFFFFFFu : (uint32_t)el;
  }
  s->applied = false;
  return BF_OK;
}
const bf_sched_run* bfh_sched_runs(const bfh_sched* s) { return s && !s->runs.empty() ? s->runs.data() : nullptr; }
int bfh_sched_tables(const bfh_sched* s, bf_sched_tables* t) {
  if (!s || !t) return BF_EINVAL;
  memset(t, 0, sizeof *t);
  t->struct_size = sizeof *t;
  t->n_stories = (uint32_t)s->story_limit.size(); t->n_queues = (uint32_t)s->queue_limit.size();
  t->global_limit = s->global_limit; t->global_running_base = s->global_base;
  t->story_limit = s->story_limit.data(); t->story_running_base = s->story_base.data();
  t->queue_limit = s->queue_limit.data(); t->queue_aging_s = s->queue_aging.data(); t->queue_running_base = s->queue_base.data();
  return BF_OK;
}
const char* bfh_sched_queue_name(const bfh_sched* s, uint32_t k) { return s && k < s->queue_names.size() ? s->queue_names[k].c_str() : ""; }

int bfh_sched_apply(bfh_sched* s) {
  if (!s) return BF_EINVAL;
  bfh_batch* b = s->batch;
  if (!b->ctx) { s->err = "batch was created without a device context"; return BF_ENODEV; }
  if (s->runs.size() != b->n) { s->err = "bfh_sched_set_run was not called for every run of the batch"; return BF_EINVAL; }
  bf_sched_tables t;
  bfh_sched_tables(s, &t);
  const uint32_t stride = BF_SCHED_STRIDE(b->L.words);
  s->records.assign((size_t)b->n * stride, 0);
  s->story_running.assign(t.n_stories, 0); s->queue_running.assign(t.n_queues, 0); s->queue_maxprio.assign(t.n_queues, 0);
  bf_sched_out out{};
  out.struct_size = sizeof out;
  out.records = s->records.data();
  out.story_running = s->story_running.data(); out.queue_running = s->queue_running.data();
  out.queue_max_priority = s->queue_maxprio.data(); out.global_running = &s->global_running;
  bf_batch bb{};
  bb.struct_size = sizeof(bf_batch);
  bb.n_runs = b->n; bb.layout = b->L;
  const int rc = bf_schedule(b->ctx, &bb, s->runs.data(), &t, &out);
  if (rc != BF_OK) { s->err = bf_last_error(b->ctx); return rc; }
  s->applied = true;
  return BF_OK;
}

int bfh_sched_steps(const bfh_sched* s, uint32_t run, int which, uint32_t* out, uint32_t cap) {
  if (!s || !s->applied || run >= s->batch->n || which < 0 || which > 2) return BF_EINVAL;
  const uint32_t W = s->batch->L.words, stride = BF_SCHED_STRIDE(W);
  const uint32_t* m = reinterpret_cast<const uint32_t*>(s->records.data() + (size_t)run * stride + 16) + (size_t)which * W;
  const uint32_t S = (uint32_t)s->batch->story_of_run[run]->flags.size();
  int n = 0;
  for (uint32_t i = 0; i < S; ++i)
    if ((m[i >> 5] >> (i & 31u)) & 1u) {
      if (out && (uint32_t)n < cap) out[n] = i;
      ++n;
    }
  return n;
}

int bfh_sched_format_message(int reason, uint32_t running, int32_t limit, char* out, size_t cap) {
  if (!out || !cap) return BF_EINVAL;
  const char* prefix;
  switch (reason) {
    case 0: prefix = "Queued due to story concurrency limit"; break;        // dag.go:104, 1797
    case BF_QUEUED_GLOBAL: prefix = "Queued due to global concurrency limit"; break;  // :106, 1850
    case BF_QUEUED_QUEUE: prefix = "Queued due to queue concurrency limit"; break;    // :105, 1852
    case BF_QUEUED_PRIORITY: snprintf(out, cap, "Queued due to higher-priority work"); return BF_OK;  // :107, 1944
    case BF_QUEUED_OTHER: snprintf(out, cap, "Queued due to scheduling limits"); return BF_OK;        // :1854
    default: out[0] = 0; return BF_OK;
  }
  snprintf(out, cap, "%s (%u running, limit %d)", prefix, running, limit);
  return BF_OK;
}

int bfh_sched_message(const bfh_sched* s, uint32_t run, int which, char* out, size_t cap) {
  if (!s || !s->applied || run >= s->batch->n || !out || !cap || (which != 1 && which != 2)) return BF_EINVAL;
  const uint32_t stride = BF_SCHED_STRIDE(s->batch->L.words);
  const bf_sched_header* h = reinterpret_cast<const bf_sched_header*>(s->records.data() + (size_t)run * stride);
  const bf_sched_run& r = s->runs[run];
  out[0] = 0;
  if (which == 1) {
    if (h->n_queued_story) return bfh_sched_format_message(0, s->story_running[r.story_key], s->story_limit[r.story_key], out, cap);
    return BF_OK;
  }
  if (!h->n_queued_sched) return BF_OK;
  if (h->sched_reason == BF_QUEUED_GLOBAL) return bfh_sched_format_message(BF_QUEUED_GLOBAL, s->global_running, s->global_limit, out, cap);
  if (h->sched_reason == BF_QUEUED_QUEUE) return bfh_sched_format_message(BF_QUEUED_QUEUE, s->queue_running[r.queue_key], s->queue_limit[r.queue_key], out, cap);
  return bfh_sched_format_message((int)h->sched_reason, 0, 0, out, cap);
}

}  // extern "C"
