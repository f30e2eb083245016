// refshape.cc — reference-SHAPED CPU restatement of one runDagIterations iteration.
//
// TEST INFRASTRUCTURE ONLY (oracle/).  Where packed_ref.c states the packed contract with bit
// loops, this file keeps the reference's own data shapes and control flow — string-keyed hash
// maps for StepStates / completed / running / dependencies, a dependency graph rebuilt on every
// pass, buildStateMaps called as often as dag.go calls it — so that (a) it is a third,
// structurally different implementation to test against and (b) its timing is an honest
// (optimistic) stand-in for what the Go reconciler does per iteration, minus Kubernetes I/O,
// JSON decoding and template evaluation.  Paths: /root/reference/internal/controller/runs/.
//
//   buildStateMaps                 dag.go:3358-3391      clearConcurrencyQueuedSteps  :2020-2051
//   buildDependencyGraphs          dag.go:3024-3073      findReadySteps               :2631-2848
//   checkSyncGates/Sleep/Wait      dag.go:1455-1547, 1217-1288, 1291-1452
//   checkSyncParallelSteps         dag.go:1112-1200      markFailFastSkipped          :3289-3312
//   markCompensationsSkipped       dag.go:3314-3342      runDagIterations body        :393-540
//
// Objects are rebuilt from the packed records (names "s<i>"), so the same seeded inputs feed
// every implementation.  Optimism, stated: the step-name regex is compiled once (Go compiles it
// per call, dag.go:3028), Phase is compared as a short string like Go's, no logger allocations.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -pthread -I../include refshape.cc -o _build/librefshape.so
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <regex>
#include <string>
#include <unordered_map>
#include <vector>

#include "bobrafrontier.h"

namespace {

using std::string;
template <class V>
using Map = std::unordered_map<string, V>;

const char* kPhaseNames[16] = {"",        "Pending", "Running",    "Succeeded", "Failed",  "Finished", "Canceled", "Compensated",
                               "Paused",  "Blocked", "Scheduling", "Timeout",   "Aborted", "Skipped",  "Pending",  ""};
const char* kQueuedMsg = "Queued due to story concurrency limit (1 running, limit 1)";
const char* kQueuedPrefixes[4] = {"Queued due to story concurrency limit", "Queued due to queue concurrency limit",
                                  "Queued due to global concurrency limit", "Queued due to higher-priority work"};

bool isTerminal(const string& p) {  // pkg/enums/enums.go:101-115
  return p == "Succeeded" || p == "Failed" || p == "Finished" || p == "Canceled" || p == "Compensated" || p == "Timeout" ||
         p == "Aborted" || p == "Skipped";
}

struct Branch { string name; bool allowFailure; };
struct Step {  // api/v1alpha1/story_types.go:156-284 (the fields the path reads)
  string name;
  std::vector<string> needs;
  int type;  // BF_STEP_*
  bool hasRef;
  bool hasIf;
  string ifExpr;
  bool allowFailure;
  bool onTimeoutSkip;
  std::vector<Branch> branches;  // parallel: with.steps
};
struct Story {
  std::vector<Step> steps, compensations, finally_;
  std::vector<Step> all() const {  // allStorySteps dag.go:3270-3280 (copies, like append does)
    std::vector<Step> v;
    v.reserve(steps.size() + compensations.size() + finally_.size());
    v.insert(v.end(), steps.begin(), steps.end());
    v.insert(v.end(), compensations.begin(), compensations.end());
    v.insert(v.end(), finally_.begin(), finally_.end());
    return v;
  }
};
struct StepState { string phase, message; };
struct StepRun { string name, stepID, phase; };
struct StoryRun {
  Map<StepState> stepStates;
  Map<std::vector<string>> primitiveChildren;
  std::vector<StepRun> stepRuns;
  Map<int> decisions;  // host-reduced gate / sleep / wait outcome per step (BF_DEC_*)
  Map<int> cond;       // host-reduced `if` outcome per step (BF_COND_*)
  bool failFast, realtime, topologyTerminated, hostGroup;
  int hostGroupValue;
};

struct StateMaps { Map<bool> completed, running, failed, allowedFailures; };

StateMaps buildStateMaps(const std::vector<Step>& steps, const Map<StepState>& states) {  // dag.go:3358-3391
  StateMaps m;
  Map<bool> allowed, allowFailure;
  for (const Step& s : steps) {
    allowed[s.name] = true;
    if (s.allowFailure) allowFailure[s.name] = true;
  }
  for (const auto& kv : states) {
    if (!allowed.count(kv.first)) continue;
    const string& ph = kv.second.phase;
    if (ph == "Succeeded" || ph == "Skipped") m.completed[kv.first] = true;
    else if (ph == "Running" || ph == "Pending" || ph == "Paused") m.running[kv.first] = true;
    else if (isTerminal(ph)) {
      if (allowFailure.count(kv.first)) { m.completed[kv.first] = true; m.allowedFailures[kv.first] = true; continue; }
      m.failed[kv.first] = true;
    }
  }
  return m;
}

bool isConcurrencyQueued(const StepState& s) {  // dag.go:2035-2051
  if (s.phase != "Pending") return false;
  for (const char* p : kQueuedPrefixes)
    if (s.message.compare(0, strlen(p), p) == 0) return true;
  return false;
}
void clearConcurrencyQueuedSteps(Map<bool>& running, const Map<StepState>& states) {  // dag.go:2020-2033
  if (running.empty() || states.empty()) return;
  for (auto it = running.begin(); it != running.end();) {
    auto st = states.find(it->first);
    if (st != states.end() && isConcurrencyQueued(st->second)) it = running.erase(it); else ++it;
  }
}

const std::regex& stepNameRegex() {  // dag.go:3028-3030 (compiled once here; per call in the reference)
  static const std::regex re(
      R"(steps\.([a-zA-Z0-9_\-]+)\.|steps\s*\[\s*['"]([a-zA-Z0-9_\-]+)['"]\s*\]|\(index\s+\.steps\s+["']([a-zA-Z0-9_\-]+)["']\))");
  return re;
}

Map<Map<bool>> buildDependencyGraphs(const std::vector<Step>& steps) {  // dag.go:3024-3073
  Map<Map<bool>> deps;
  Map<Map<bool>> dependents;
  for (const Step& s : steps) {
    deps[s.name];
    dependents[s.name];
    for (const string& d : s.needs) { deps[s.name][d] = true; dependents[d][s.name] = true; }
    if (s.hasIf) {
      auto b = std::sregex_iterator(s.ifExpr.begin(), s.ifExpr.end(), stepNameRegex());
      for (auto it = b; it != std::sregex_iterator(); ++it) {
        const std::smatch& m = *it;
        string d = m[3].matched ? m[3].str() : m[2].matched ? m[2].str() : m[1].str();
        if (!d.empty()) { deps[s.name][d] = true; dependents[d][s.name] = true; }
      }
    }
  }
  return deps;
}

bool depSatisfiedForRealtime(bool realtime, const StepState& ds) {  // dag.go:3457-3473
  if (!realtime) return false;
  if (ds.phase.empty()) return false;
  if (isTerminal(ds.phase)) return ds.phase == "Succeeded";
  return ds.phase == "Pending" || ds.phase == "Running" || ds.phase == "Paused";
}

struct ReadyOut { std::vector<string> ready, skipped, failedNow, evaluatedIf; Map<string> reasons; };

// findReadySteps dag.go:2631-2848 with skip_max (a failed dep found before any unmet one)
ReadyOut findReadySteps(StoryRun& run, const std::vector<Step>& steps, Map<bool>& completed, Map<bool>& running,
                        Map<Map<bool>>& dependencies, bool allowFailed, bool skipOnFailed) {
  ReadyOut out;
  for (const Step& step : steps) {
    if (completed.count(step.name) || running.count(step.name)) continue;
    auto sit = run.stepStates.find(step.name);
    if (sit != run.stepStates.end() && isTerminal(sit->second.phase)) continue;
    bool allDepsMet = true;
    string failedDep;
    const Map<bool>& deps = dependencies[step.name];
    for (int pass = 0; pass < 2 && failedDep.empty(); ++pass) {  // pass 0: look for a failed dep (skip_max), pass 1: unmet
      for (const auto& kv : deps) {
        const string& dep = kv.first;
        if (completed.count(dep)) continue;
        auto dit = run.stepStates.find(dep);
        const StepState ds = dit == run.stepStates.end() ? StepState{} : dit->second;
        if (depSatisfiedForRealtime(run.realtime, ds)) continue;
        if (allowFailed && isTerminal(ds.phase)) continue;
        if (skipOnFailed && isTerminal(ds.phase) && ds.phase != "Succeeded" && ds.phase != "Skipped") {
          if (pass == 0) { failedDep = dep; break; }
          continue;
        }
        if (pass == 1) { allDepsMet = false; break; }
      }
    }
    if (!failedDep.empty()) {
      out.skipped.push_back(step.name);
      out.reasons[step.name] = "Skipped due to failed dependency: " + failedDep;
      continue;
    }
    if (!allDepsMet) continue;
    if (step.hasIf && !run.realtime) out.evaluatedIf.push_back(step.name);
    auto cit = run.cond.find(step.name);
    const int code = cit == run.cond.end() ? BF_COND_PASS : cit->second;
    if (code == BF_COND_FAIL) {  // dag.go:2744-2748 / 2810-2814
      run.stepStates[step.name] = StepState{"Failed", "template expression uses disallowed function 'env'"};
      out.failedNow.push_back(step.name);
      continue;
    }
    if (code == BF_COND_HOLD) continue;
    if (code == BF_COND_SKIP) {
      out.skipped.push_back(step.name);
      out.reasons[step.name] = "Skipped due to 'if' condition";
      continue;
    }
    out.ready.push_back(step.name);
  }
  return out;
}

// gate / sleep / wait share the decision mapping once time and `until` are reduced on the host
bool checkSyncPrimitives(StoryRun& run, const std::vector<Step>& steps) {  // dag.go:1455-1547, 1217-1288, 1291-1452
  bool updated = false;
  for (const Step& step : steps) {
    if (step.hasRef || (step.type != BF_STEP_GATE && step.type != BF_STEP_SLEEP && step.type != BF_STEP_WAIT)) continue;
    auto it = run.stepStates.find(step.name);
    if (it == run.stepStates.end() || isTerminal(it->second.phase)) continue;
    const string& cur = it->second.phase;
    if (cur != "Paused" && cur != "Running" && cur != "Pending") continue;
    auto dit = run.decisions.find(step.name);
    const int d = dit == run.decisions.end() ? BF_DEC_PENDING : dit->second;
    string next;
    switch (d) {
      case BF_DEC_SUCCEED: next = "Succeeded"; break;
      case BF_DEC_FAIL: next = "Failed"; break;
      case BF_DEC_TIMED_OUT: next = step.onTimeoutSkip ? "Skipped" : "Timeout"; break;
      default: next = "Paused"; break;
    }
    if (next != cur) { it->second.phase = next; it->second.message = "synced"; updated = true; }
  }
  return updated;
}

bool checkSyncParallelSteps(StoryRun& run, const std::vector<Step>& steps) {  // dag.go:1112-1200
  if (run.primitiveChildren.empty()) return false;
  Map<const StepRun*> byName;
  for (const StepRun& sr : run.stepRuns) byName[sr.name] = &sr;
  bool updated = false;
  for (const Step& step : steps) {
    if (step.hasRef || step.type != BF_STEP_PARALLEL) continue;
    auto it = run.stepStates.find(step.name);
    if (it == run.stepStates.end() || isTerminal(it->second.phase)) continue;
    auto cit = run.primitiveChildren.find(step.name);
    if (cit == run.primitiveChildren.end() || cit->second.empty()) continue;
    Map<bool> allowFailure;
    for (const Branch& b : step.branches)
      if (b.allowFailure) allowFailure[b.name] = true;
    bool allDone = true;
    std::vector<string> failedBranches;
    for (const string& childName : cit->second) {
      auto c = byName.find(childName);
      if (c == byName.end() || c->second->phase.empty() || !isTerminal(c->second->phase)) { allDone = false; continue; }
      if (c->second->phase == "Succeeded" || c->second->phase == "Skipped") continue;
      if (allowFailure.count(c->second->stepID)) continue;
      failedBranches.push_back(c->second->stepID);
    }
    if (!allDone) continue;
    it->second.phase = failedBranches.empty() ? "Succeeded" : "Failed";
    updated = true;
  }
  return updated;
}

bool stepsTerminal(size_t total, const Map<bool>& completed, const Map<bool>& failed) {  // dag.go:3282-3287
  return total == 0 || completed.size() + failed.size() == total;
}

struct IterOut { int group; uint32_t sum; ReadyOut ready; };

IterOut runIteration(const Story& story, StoryRun& run) {  // the loop body, dag.go:393-540 (no launch effects)
  IterOut o{};
  const std::vector<Step> allSteps = story.all();
  checkSyncPrimitives(run, allSteps);       // :409-415  (tier K1 skips stage I only)
  checkSyncParallelSteps(run, allSteps);    // :418
  if (!run.hostGroup) {
    StateMaps mm = buildStateMaps(story.steps, run.stepStates);  // :422
    clearConcurrencyQueuedSteps(mm.running, run.stepStates);
    if (run.failFast && !mm.failed.empty()) {  // :424-430, markFailFastSkipped :3289-3312
      for (const Step& s : story.steps) {
        if (mm.completed.count(s.name) || mm.running.count(s.name)) continue;
        StepState st = run.stepStates.count(s.name) ? run.stepStates[s.name] : StepState{};
        if (isTerminal(st.phase)) continue;
        run.stepStates[s.name] = StepState{"Skipped", "Skipped due to fail-fast policy"};
      }
      mm = buildStateMaps(story.steps, run.stepStates);
      clearConcurrencyQueuedSteps(mm.running, run.stepStates);
    }
    bool mainDone = stepsTerminal(story.steps.size(), mm.completed, mm.failed);  // :431
    if (!mainDone && run.realtime && run.topologyTerminated) {                    // :436-464
      mainDone = true;
      for (const Step& s : story.steps) {
        auto it = run.stepStates.find(s.name);
        if (it != run.stepStates.end() && !isTerminal(it->second.phase)) { it->second.phase = "Failed"; it->second.message = "realtime topology terminated"; }
      }
      mm = buildStateMaps(story.steps, run.stepStates);
      clearConcurrencyQueuedSteps(mm.running, run.stepStates);
    }
    if (mainDone && mm.failed.empty() && !story.compensations.empty()) {  // :466, markCompensationsSkipped :3314-3342
      StateMaps cm = buildStateMaps(story.compensations, run.stepStates);
      for (const Step& s : story.compensations) {
        if (cm.completed.count(s.name) || cm.running.count(s.name) || cm.failed.count(s.name)) continue;
        StepState st = run.stepStates.count(s.name) ? run.stepStates[s.name] : StepState{};
        if (isTerminal(st.phase)) continue;
        run.stepStates[s.name] = StepState{"Skipped", "Skipped because story succeeded"};
      }
    }
    StateMaps cm = buildStateMaps(story.compensations, run.stepStates);  // :472
    StateMaps fm = buildStateMaps(story.finally_, run.stepStates);       // :473
    const bool compDone = stepsTerminal(story.compensations.size(), cm.completed, cm.failed);
    const bool finalDone = stepsTerminal(story.finally_.size(), fm.completed, fm.failed);
    if (!mainDone) o.group = BF_GROUP_MAIN;                                                            // :484
    else if (!mm.failed.empty() && !story.compensations.empty() && !compDone) o.group = BF_GROUP_COMPENSATION;
    else if (!story.finally_.empty() && !finalDone) o.group = BF_GROUP_FINALLY;
    else o.group = BF_GROUP_DONE;
    o.sum = (mainDone ? BF_SUM_MAIN_DONE : 0u) | (!mm.failed.empty() ? BF_SUM_MAIN_FAILED : 0u) | (compDone ? BF_SUM_COMP_DONE : 0u) |
            (finalDone ? BF_SUM_FINAL_DONE : 0u) | (!cm.failed.empty() ? BF_SUM_COMP_FAILED : 0u) | (!fm.failed.empty() ? BF_SUM_FINAL_FAILED : 0u);
  } else {
    o.group = run.hostGroupValue;
  }
  if (o.group == BF_GROUP_DONE) return o;
  StateMaps am = buildStateMaps(allSteps, run.stepStates);  // :497
  clearConcurrencyQueuedSteps(am.running, run.stepStates);
  const bool allowFailed = o.group != BF_GROUP_MAIN;                    // :500
  const bool skipOnFailed = o.group == BF_GROUP_MAIN && !run.failFast;  // :501
  const std::vector<Step>& list = o.group == BF_GROUP_COMPENSATION ? story.compensations : o.group == BF_GROUP_FINALLY ? story.finally_ : story.steps;
  Map<Map<bool>> deps = buildDependencyGraphs(list);  // :1700
  o.ready = findReadySteps(run, list, am.completed, am.running, deps, allowFailed, skipOnFailed);
  return o;
}

// ---------------------------------------------------------------------------------- packed <-> objects
struct Topo {
  uint32_t n_steps, n_edges;
  const uint32_t* row_ptr; const uint16_t* col_idx; const uint8_t* step_flags;
  const bf_parallel_desc* parallel; uint32_t n_parallel; uint32_t pad;
  const uint8_t* branch_allow_bits; const uint32_t* child_first;
};

inline int getCode(const uint8_t* base, uint32_t W, int nbits, uint32_t i) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(base);
  int v = 0;
  for (int b = 0; b < nbits; ++b) v |= (int)((w[(uint32_t)b * W + (i >> 5)] >> (i & 31u)) & 1u) << b;
  return v;
}
inline void setCode(uint8_t* base, uint32_t W, int nbits, uint32_t i, int v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(base);
  for (int b = 0; b < nbits; ++b)
    if ((v >> b) & 1) w[(uint32_t)b * W + (i >> 5)] |= 1u << (i & 31u);
}

struct Batch {
  bf_layout L;
  std::vector<std::shared_ptr<Story>> storyOfSlot;
  std::vector<std::shared_ptr<Story>> storyOfRun;
  std::vector<StoryRun> runs;
  std::vector<std::vector<string>> namesOfRun;  // index -> name (shared per slot in practice)
};

string stepName(uint32_t i) { return "s" + std::to_string(i); }

std::shared_ptr<Story> storyFromTopo(const Topo& T) {
  auto st = std::make_shared<Story>();
  uint32_t q = 0;
  for (uint32_t i = 0; i < T.n_steps; ++i) {
    const uint8_t f = T.step_flags[i];
    Step s;
    s.name = stepName(i);
    for (uint32_t e = T.row_ptr[i]; e < T.row_ptr[i + 1]; ++e) s.needs.push_back(stepName(T.col_idx[e]));
    s.type = f & BF_SF_TYPE_MASK;
    s.hasRef = s.type == BF_STEP_ENGRAM;
    s.hasIf = (f & BF_SF_HAS_IF) != 0;
    if (s.hasIf) s.ifExpr = "{{ inputs.c_" + s.name + " }}";
    s.allowFailure = (f & BF_SF_ALLOW_FAILURE) != 0;
    s.onTimeoutSkip = (f & BF_SF_ON_TIMEOUT_SKIP) != 0;
    if (s.type == BF_STEP_PARALLEL && q < T.n_parallel && T.parallel[q].step == i) {
      for (uint32_t b = 0; b < T.parallel[q].branches; ++b) {
        const uint32_t bit = T.parallel[q].allow_first + b;
        const bool al = T.branch_allow_bits && ((T.branch_allow_bits[bit >> 3] >> (bit & 7u)) & 1u);
        s.branches.push_back(Branch{"b" + std::to_string(b), al});
      }
      ++q;
    }
    const int g = (f & BF_SF_GROUP_MASK) >> BF_SF_GROUP_SHIFT;
    (g == BF_GROUP_MAIN ? st->steps : g == BF_GROUP_COMPENSATION ? st->compensations : st->finally_).push_back(std::move(s));
  }
  return st;
}

}  // namespace

extern "C" {

// Build object-form Stories / StoryRuns from packed records (untimed preparation).
void* orc_refshape_build(const void* topos_v, uint32_t n_topos, const bf_layout* L, uint32_t n_runs, const uint8_t* state) {
  const Topo* topos = static_cast<const Topo*>(topos_v);
  auto* B = new Batch();
  B->L = *L;
  B->storyOfSlot.resize(n_topos);
  B->runs.resize(n_runs);
  B->storyOfRun.resize(n_runs);
  const uint32_t W = L->words;
  for (uint32_t r = 0; r < n_runs; ++r) {
    const uint8_t* rec = state + (size_t)r * L->state_stride;
    const bf_run_header* h = reinterpret_cast<const bf_run_header*>(rec);
    if (h->topo_slot >= n_topos || topos[h->topo_slot].n_steps == 0) { delete B; return nullptr; }
    const Topo& T = topos[h->topo_slot];
    if (!B->storyOfSlot[h->topo_slot]) B->storyOfSlot[h->topo_slot] = storyFromTopo(T);
    B->storyOfRun[r] = B->storyOfSlot[h->topo_slot];
    StoryRun& run = B->runs[r];
    run.failFast = h->run_flags & BF_RF_FAIL_FAST;
    run.realtime = h->run_flags & BF_RF_REALTIME;
    run.topologyTerminated = h->run_flags & BF_RF_TOPOLOGY_TERMINATED;
    run.hostGroup = h->run_flags & BF_RF_HOST_GROUP;
    run.hostGroupValue = (h->run_flags >> BF_RF_HOST_GROUP_SHIFT) & 3;
    uint32_t q = 0;
    for (uint32_t i = 0; i < T.n_steps; ++i) {
      const string name = stepName(i);
      const int p = getCode(rec + L->off_phase, W, 4, i);
      if (p != 0 && p != 15) run.stepStates[name] = StepState{kPhaseNames[p], p == BF_PHASE_PENDING_QUEUED ? kQueuedMsg : ""};
      if (L->off_cond != BF_OFF_NONE) { const int c = getCode(rec + L->off_cond, W, 2, i); if (c) run.cond[name] = c; }
      if (L->off_decision != BF_OFF_NONE) { const int d = getCode(rec + L->off_decision, W, 2, i); if (d) run.decisions[name] = d; }
      const int ty = T.step_flags[i] & BF_SF_TYPE_MASK;
      if (ty == BF_STEP_PARALLEL && q < T.n_parallel && T.parallel[q].step == i) {
        if (L->off_child != BF_OFF_NONE && ((h->children_registered >> q) & 1ull)) {
          std::vector<string> kids;
          for (uint32_t b = 0; b < T.parallel[q].branches; ++b) {
            const string cname = name + "-b" + std::to_string(b);
            kids.push_back(cname);
            const uint32_t ci = T.child_first[q] + b;
            int cp = (rec[L->off_child + (ci >> 1)] >> ((ci & 1u) * 4u)) & 0xF;
            if (cp == 15) cp = 0;
            run.stepRuns.push_back(StepRun{cname, "b" + std::to_string(b), kPhaseNames[cp]});
          }
          run.primitiveChildren[name] = kids;
        }
        ++q;
      }
    }
  }
  return B;
}

void orc_refshape_free(void* h) { delete static_cast<Batch*>(h); }

struct RsJob { Batch* B; uint8_t* result; uint32_t lo, hi; uint64_t evals; };

static void* rsWorker(void* arg) {
  RsJob* j = static_cast<RsJob*>(arg);
  const bf_layout& L = j->B->L;
  const uint32_t W = L.words;
  j->evals = 0;
  for (uint32_t r = j->lo; r < j->hi; ++r) {
    const Story& story = *j->B->storyOfRun[r];
    StoryRun run = j->B->runs[r];  // the reconciler works on a DeepCopy from the informer cache
    Map<string> before;
    for (const auto& kv : run.stepStates) before[kv.first] = kv.second.phase;
    IterOut o = runIteration(story, run);
    uint8_t* rec = j->result + (size_t)r * L.result_stride;
    memset(rec, 0, L.result_stride);
    bf_result_header* oh = reinterpret_cast<bf_result_header*>(rec);
    auto idx = [](const string& n) { return (uint32_t)std::stoul(n.substr(1)); };
    auto setBit = [&](uint32_t off, const string& n) { if (off != BF_OFF_NONE) { uint32_t i = idx(n); reinterpret_cast<uint32_t*>(rec + off)[i >> 5] |= 1u << (i & 31u); } };
    for (const string& n : o.ready.ready) setBit(L.off_ready, n);
    for (const string& n : o.ready.skipped) {
      setBit(L.off_skip, n);
      if (o.ready.reasons[n].find("failed dependency") != string::npos) setBit(L.off_skip_dep, n);
    }
    for (const string& n : o.ready.failedNow) setBit(L.off_fail, n);
    for (const string& n : o.ready.evaluatedIf) setBit(L.off_needs_cond, n);
    uint32_t nexp = 0;
    const size_t S = story.steps.size() + story.compensations.size() + story.finally_.size();
    bool changed = false;
    const std::vector<Step> all = story.all();
    for (const Step& s : all) {
      auto it = run.stepStates.find(s.name);
      int code = 0;
      if (it != run.stepStates.end()) {
        for (int c = 1; c <= 13; ++c)
          if (it->second.phase == kPhaseNames[c]) code = c;
        if (isConcurrencyQueued(it->second)) code = BF_PHASE_PENDING_QUEUED;
      }
      if (L.off_phase_out != BF_OFF_NONE) setCode(rec + L.off_phase_out, W, 4, idx(s.name), code);
      auto b = before.find(s.name);
      const string was = b == before.end() ? "" : b->second;
      const string now = it == run.stepStates.end() ? "" : it->second.phase;
      if (was != now) changed = true;
      if (s.type == BF_STEP_PARALLEL && !s.hasRef && std::find(o.ready.ready.begin(), o.ready.ready.end(), s.name) != o.ready.ready.end())
        nexp += (uint32_t)s.branches.size();
    }
    oh->summary = o.sum | (uint32_t)o.group | (changed ? BF_SUM_PHASE_CHANGED : 0u) | (1u << BF_SUM_ITER_SHIFT);
    oh->n_ready = (uint32_t)o.ready.ready.size();
    oh->n_skip = (uint32_t)o.ready.skipped.size();
    oh->n_expansion = nexp;
    j->evals += S;
  }
  return nullptr;
}

// One runDagIterations iteration per StoryRun, `threads` workers (the reference runs 8 reconcile workers by
// default: internal/config/controller_config.go:721).  Returns total evals.
uint64_t orc_refshape_run(void* h, uint8_t* result, int threads) {
  Batch* B = static_cast<Batch*>(h);
  const uint32_t n = (uint32_t)B->runs.size();
  if (threads < 1) threads = 1;
  if ((uint32_t)threads > n) threads = n ? (int)n : 1;
  std::vector<RsJob> jobs(threads);
  std::vector<pthread_t> th(threads);
  const uint32_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    jobs[t] = RsJob{B, result, std::min(n, (uint32_t)t * per), std::min(n, ((uint32_t)t + 1) * per), 0};
    if (t) pthread_create(&th[t], nullptr, rsWorker, &jobs[t]);
  }
  rsWorker(&jobs[0]);
  uint64_t evals = jobs[0].evals;
  for (int t = 1; t < threads; ++t) { pthread_join(th[t], nullptr); evals += jobs[t].evals; }
  return evals;
}

}  // extern "C"
