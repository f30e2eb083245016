"""Object form -> packed records (test-side packer).

Turns the CRD-shaped objects of oracle/pyoracle.py (Story / StoryRun / StepRun) into the
records of include/bobrafrontier.h, doing the host work the contract assigns to the packer:
name->index maps, regex/alias dependency extraction (dag.go:3024-3073, via the oracle's own
build_dependency_graphs), queued-message folding (dag.go:2035-2051), gate/sleep/wait decision
codes (time and `until` evaluation are host work) and `if` condition codes.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bobrapet_b200 import _abi as A  # noqa: E402
from bobrapet_b200.frontier import TopologySet  # noqa: E402
from bobrapet_b200.records import PAR_DTYPE, make_layout, pack_state  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from oracle.packed import child_first_of  # noqa: E402


@dataclass
class PackedStory:
    names: List[str]
    index: Dict[str, int]
    row_ptr: np.ndarray
    col_idx: np.ndarray
    flags: np.ndarray
    par_steps: List[int] = field(default_factory=list)
    par_branches: List[List[str]] = field(default_factory=list)
    par_allow: List[List[bool]] = field(default_factory=list)

    @property
    def S(self):
        return len(self.names)

    def child_first(self):
        return child_first_of([len(b) for b in self.par_branches])

    def child_nibbles(self):
        cf = self.child_first()
        if not len(cf):
            return 0
        return int((int(cf[-1]) + len(self.par_branches[-1]) + 7) // 8 * 8)


def pack_story(story: O.Story) -> PackedStory:
    groups = [(story.steps, A.GROUP_MAIN), (story.compensations, A.GROUP_COMPENSATION), (story.finally_, A.GROUP_FINALLY)]
    names = [s.name for s in O.all_story_steps(story)]
    index = {n: i for i, n in enumerate(names)}
    assert len(index) == len(names), "duplicate step names"
    rows: List[List[int]] = []
    flags = []
    ps = PackedStory(names, index, None, None, None)
    for steps, g in groups:
        deps, _ = O.build_dependency_graphs(steps)  # built per evaluated list, dag.go:1700
        for st in steps:
            row = []
            for d in deps.get(st.name, {}):
                if d not in index:
                    raise ValueError("dangling dependency %s->%s (unsupported by the packer)" % (st.name, d))
                row.append(index[d])
            rows.append(sorted(set(row)))
            f = A.STEP_ENGRAM if st.ref else A.STEP_TYPE_CODE[st.type]
            if st.allow_failure:
                f |= A.SF_ALLOW_FAILURE
            if st.if_:
                f |= A.SF_HAS_IF
            if st.type in ("gate", "wait") and isinstance(st.with_, dict) and str(st.with_.get("onTimeout", "")).strip().lower() == "skip":
                f |= A.SF_ON_TIMEOUT_SKIP
            f |= g << A.SF_GROUP_SHIFT
            flags.append(f)
            if st.type == "parallel" and not st.ref:
                br = O.parse_parallel_branches(st) if st.with_ is not None else []
                ps.par_steps.append(index[st.name])
                ps.par_branches.append([b.name for b in br])
                ps.par_allow.append([bool(b.allow_failure) for b in br])
    rp = [0]
    for r in rows:
        rp.append(rp[-1] + len(r))
    ps.row_ptr = np.asarray(rp, dtype=np.uint32)
    ps.col_idx = np.asarray([c for r in rows for c in r], dtype=np.uint16)
    ps.flags = np.asarray(flags, dtype=np.uint8)
    return ps


def topology_set(stories: List[PackedStory]) -> TopologySet:
    S = [p.S for p in stories]
    E = [len(p.col_idx) for p in stories]
    P = [len(p.par_steps) for p in stories]
    par = np.zeros(sum(P), dtype=PAR_DTYPE)
    bits: List[bool] = []
    k = 0
    for p in stories:
        for stp, br, al in zip(p.par_steps, p.par_branches, p.par_allow):
            par[k] = (stp, len(br), len(bits))
            bits.extend(al)
            k += 1
    allow = np.packbits(np.asarray(bits + [False] * ((-len(bits)) % 8), dtype=bool), bitorder="little") if bits else None
    return TopologySet(S, E, np.concatenate([p.row_ptr for p in stories]),
                       np.concatenate([p.col_idx for p in stories]) if sum(E) else np.zeros(0, np.uint16),
                       np.concatenate([p.flags for p in stories]), P, par, allow)


def phase_code(st: Optional[O.StepState]) -> int:
    if st is None or st.phase == "":
        return A.PHASE_NONE
    if O.is_concurrency_queued(st):
        return A.PHASE_PENDING_QUEUED
    return O.PHASE_CODE[st.phase]


def decision_code(step: O.Step, srun: O.StoryRun, now: float, timers: Optional[O.StepTimers],
                  evaluator=None, vars_=None, offloaded_policy="fail") -> int:
    """Host-side reduction of gate / sleep / wait to a 2-bit decision (time + templates stay on the host)."""
    cur = srun.step_states.get(step.name)
    started = cur.started_at if (cur is not None and cur.started_at is not None) else now
    if step.type == "gate":
        try:
            timeout, _p, _o = O.parse_gate_config(step)
        except ValueError:
            return A.DEC_FAIL
        gs = srun.gates.get(step.name)
        if gs is not None and gs.state == "Approved":
            return A.DEC_SUCCEED
        if gs is not None and gs.state == "Rejected":
            return A.DEC_FAIL
        if timeout is not None:
            at = started + timeout
            if timers is not None and step.name in timers.gate_timeout_at:
                at = timers.gate_timeout_at[step.name]
            if not (now < at):
                return A.DEC_TIMED_OUT
        return A.DEC_PENDING
    if step.type == "sleep":
        try:
            dur = O.parse_sleep_config(step)
        except ValueError:
            return A.DEC_FAIL
        if dur <= 0:
            return A.DEC_SUCCEED
        until = started + dur
        if timers is not None and step.name in timers.sleep_until:
            until = timers.sleep_until[step.name]
        return A.DEC_SUCCEED if until - now <= 0 else A.DEC_PENDING
    if step.type == "wait":
        try:
            until_expr, timeout, _p, _o = O.parse_wait_config(step)
        except ValueError:
            return A.DEC_FAIL
        if O.validate_template_string(until_expr) is not None:
            return A.DEC_FAIL
        result = False
        try:
            result = evaluator(step.name, until_expr, vars_ or {}) if evaluator else False
        except O.EvaluationBlocked:
            result = False
        except O.OffloadedDataUsage:
            if offloaded_policy != "block":
                return A.DEC_FAIL
        except Exception:
            result = False
        if result:
            return A.DEC_SUCCEED
        if timeout is not None:
            at = started + timeout
            if timers is not None and step.name in timers.wait_timeout_at:
                at = timers.wait_timeout_at[step.name]
            if not (now < at):
                return A.DEC_TIMED_OUT
        return A.DEC_PENDING
    return A.DEC_PENDING


def cond_code(step: O.Step, story: O.Story, evaluator, vars_, stale=None, offloaded_policy="fail") -> int:
    """Host-side reduction of the `if` / stale-`with` outcome to BF_COND_* (dag.go:2741-2843)."""
    stale = stale or (lambda _s, _k: False)
    if step.if_ and not story.realtime:
        if O.validate_template_string(step.if_) is not None:
            return A.COND_FAIL
        try:
            result = evaluator(step.name, step.if_, vars_ or {}) if evaluator else False
        except O.EvaluationBlocked:
            return A.COND_HOLD
        except O.OffloadedDataUsage:
            return A.COND_HOLD if offloaded_policy in ("block", "controller", "inject") else A.COND_FAIL
        except Exception:
            return A.COND_HOLD
        if not result:
            return A.COND_HOLD if stale(step.name, "if") else A.COND_SKIP
    if stale(step.name, "with"):
        return A.COND_HOLD
    return A.COND_PASS


def run_flags_of(story: O.Story, srun: O.StoryRun, host_group: Optional[str] = None) -> int:
    f = 0
    if O.should_fail_fast(story):
        f |= A.RF_FAIL_FAST
    if story.realtime:
        f |= A.RF_REALTIME
    if srun.topology_terminated:
        f |= A.RF_TOPOLOGY_TERMINATED
    if host_group is not None:
        g = {"main": 0, "compensation": 1, "finally": 2, "finalize": 3}[host_group]
        f |= A.RF_HOST_GROUP | (g << A.RF_HOST_GROUP_SHIFT)
    return f


def pack_runs(stories: List[O.Story], packed: List[PackedStory], sruns: List[O.StoryRun], story_of_run: List[int],
              slots, step_runs: Optional[List[Optional[List[O.StepRun]]]] = None, evaluator=None, vars_=None,
              now: float = 0.0, timers: Optional[List[Optional[O.StepTimers]]] = None, stale=None,
              host_groups: Optional[List[Optional[str]]] = None, fields: int = A.F_COND | A.F_DECISION | A.F_ALL_OUT,
              offloaded_policy="fail"):
    """-> (layout, state records).  One record per StoryRun."""
    n = len(sruns)
    s_max = max(p.S for p in packed)
    child_max = max(p.child_nibbles() for p in packed)
    if child_max:
        fields |= A.F_CHILD
    L = make_layout(s_max, child_max, fields)
    phase = np.zeros((n, s_max), np.uint8)
    cond = np.zeros((n, s_max), np.uint8)
    dec = np.zeros((n, s_max), np.uint8)
    child = np.zeros((n, max(child_max, 1)), np.uint8)
    reg = np.zeros(n, np.uint64)
    rflags = np.zeros(n, np.uint8)
    sl = np.zeros(n, np.uint32)
    for r, srun in enumerate(sruns):
        story, ps = stories[story_of_run[r]], packed[story_of_run[r]]
        sl[r] = slots[story_of_run[r]]
        rflags[r] = run_flags_of(story, srun, host_groups[r] if host_groups else None)
        tm = timers[r] if timers else None
        for i, st in enumerate(O.all_story_steps(story)):
            phase[r, i] = phase_code(srun.step_states.get(st.name))
            cond[r, i] = cond_code(st, story, evaluator, vars_, stale, offloaded_policy)
            if st.type in ("gate", "sleep", "wait") and not st.ref:
                dec[r, i] = decision_code(st, srun, now, tm, evaluator, vars_, offloaded_policy)
        srs = {sr.name: sr for sr in (step_runs[r] or [])} if step_runs else {}
        cf = ps.child_first()
        for q, stp in enumerate(ps.par_steps):
            pname = ps.names[stp]
            kids = srun.primitive_children.get(pname) or []
            if step_runs is None or step_runs[r] is None or not kids:
                continue
            reg[r] |= np.uint64(1) << np.uint64(q)
            # children are matched to branches by StepID (dag.go:1172); order = branch order
            by_id = {}
            for kn in kids:
                sr = srs.get(kn)
                if sr is not None:
                    by_id[sr.step_id] = sr
            for b, bname in enumerate(ps.par_branches[q]):
                sr = by_id.get(bname)
                child[r, int(cf[q]) + b] = O.PHASE_CODE.get(sr.phase, 0) if sr is not None else 0
    state = pack_state(L, sl, rflags, phase, cond, dec, child if child_max else None, reg)
    return L, state
