"""pyoracle — object-shaped CPU restatement of the bobrapet DAG frontier path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may
import it, and only as the checker.

Why a restatement: the reference is Go (go 1.26.2, k8s.io/* v0.35.4,
controller-runtime v0.23.3, github.com/bubustack/core v0.1.5 un-vendored) and
there is no Go toolchain in this image, so the reference cannot be compiled or
run here.  This file restates, function by function and with the reference's
own data shapes (string-keyed dicts standing in for Go maps), the functions on
the hot path.  Every function cites the reference lines it follows
(paths relative to /root/reference).  It is pinned by the reference's own
known-answer tests (internal/controller/runs/dag_test.go), re-expressed in
tests/test_oracle_kat.py.

Third-party arithmetic: the truth value of `if` / `until` expressions comes
from templating.Evaluator.EvaluateCondition in github.com/bubustack/core v0.1.5
(call sites dag.go:1373, 2679, 2771).  That module is absent; the oracle takes
an `evaluator` callback instead and the kernel takes the RESULT as an input
code, so expression evaluation itself is out of the parity claim.

Reference nondeterminism (Go map iteration order, dag.go:2714-2733) is made
explicit: find_ready_steps() takes a `dep_order` policy and the packed contract
pins skip_max.
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple

# --------------------------------------------------------------------------
# pkg/enums/enums.go:35-115 — Phase, IsTerminal; :143-174 StepType
# --------------------------------------------------------------------------
PHASES = [
    "", "Pending", "Running", "Succeeded", "Failed", "Finished", "Canceled",
    "Compensated", "Paused", "Blocked", "Scheduling", "Timeout", "Aborted", "Skipped",
]
PHASE_CODE = {p: i for i, p in enumerate(PHASES)}
_TERMINAL = {"Succeeded", "Failed", "Finished", "Canceled", "Compensated", "Timeout", "Aborted", "Skipped"}


def is_terminal(phase: str) -> bool:
    """Phase.IsTerminal, pkg/enums/enums.go:101-115."""
    return phase in _TERMINAL


STEP_TYPES = ["", "condition", "parallel", "sleep", "stop", "wait", "executeStory", "gate"]

# dag.go:103-108
QUEUED_PREFIXES = (
    "Queued due to story concurrency limit",
    "Queued due to queue concurrency limit",
    "Queued due to global concurrency limit",
    "Queued due to higher-priority work",
)
SLEEP_COMPLETED_MSG = "Sleep completed."


# --------------------------------------------------------------------------
# api/v1alpha1/story_types.go:156-284 (Step), :90-150 (StorySpec)
# api/runs/v1alpha1/storyrun_types.go:106-287 (StoryRunStatus, StepState, GateStatus)
# --------------------------------------------------------------------------
@dataclass
class Step:
    name: str
    needs: List[str] = field(default_factory=list)
    type: str = ""                 # StepType; "" with ref=True means engram
    ref: bool = False              # step.Ref != nil
    if_: Optional[str] = None      # step.If
    allow_failure: Optional[bool] = None
    with_: Optional[Any] = None    # step.With (decoded JSON object) or None

    def with_raw(self) -> str:
        return "" if self.with_ is None else json.dumps(self.with_)


@dataclass
class Story:
    steps: List[Step] = field(default_factory=list)
    compensations: List[Step] = field(default_factory=list)
    finally_: List[Step] = field(default_factory=list)
    realtime: bool = False                              # Spec.Pattern.IsRealtime(), enums.go:335
    continue_on_step_failure: Optional[bool] = None     # Policy.Retries.ContinueOnStepFailure


@dataclass
class StepState:
    phase: str = ""
    message: str = ""
    started_at: Optional[float] = None
    finished_at: Optional[float] = None

    def copy(self) -> "StepState":
        return StepState(self.phase, self.message, self.started_at, self.finished_at)


@dataclass
class GateStatus:
    state: str = ""       # "", "Pending", "Approved", "Rejected" (storyrun_types.go:265-272)
    message: str = ""


@dataclass
class StepRun:
    name: str
    step_id: str
    phase: str = ""


@dataclass
class StoryRun:
    step_states: Dict[str, StepState] = field(default_factory=dict)
    gates: Dict[str, GateStatus] = field(default_factory=dict)
    primitive_children: Dict[str, List[str]] = field(default_factory=dict)
    topology_terminated: bool = False   # Degraded=True, reason TopologyTerminated (dag.go:437-441)
    allowed_failures: List[str] = field(default_factory=list)
    phase: str = ""                     # StoryRun.Status.Phase (set by stop steps)


@dataclass
class StepTimers:
    """stepTimerStore items, dag.go:110-171."""
    sleep_until: Dict[str, float] = field(default_factory=dict)
    wait_timeout_at: Dict[str, float] = field(default_factory=dict)
    gate_timeout_at: Dict[str, float] = field(default_factory=dict)


@dataclass
class DepPolicy:
    """stepDependencyPolicy, dag.go:98-101."""
    allow_failed_dependencies: bool = False
    skip_on_failed_dependency: bool = False


class EvaluationBlocked(Exception):
    """templating.ErrEvaluationBlocked."""


class OffloadedDataUsage(Exception):
    """templating.ErrOffloadedDataUsage."""


# evaluator(step_name, expr, vars) -> bool ; may raise EvaluationBlocked /
# OffloadedDataUsage / Exception.  Stands in for templating.Evaluator.
Evaluator = Callable[[str, str, Dict[str, Any]], bool]


def literal_evaluator(step_name: str, expr: str, vars_: Dict[str, Any]) -> bool:
    """Minimal stand-in good for the reference's own three pinned cases:
    literal "true"/"false" (dag_test.go:699) and `{{ inputs.X }}` (:1283,:1322)."""
    e = expr.strip()
    m = re.fullmatch(r"\{\{\s*inputs\.([A-Za-z0-9_]+)\s*\}\}", e)
    if m:
        return bool((vars_.get("inputs") or {}).get(m.group(1), False))
    if e == "true":
        return True
    if e == "false":
        return False
    raise Exception("unsupported expression in literal_evaluator: %r" % expr)


# --------------------------------------------------------------------------
# pkg/templatesafety/templatesafety.go:21-80
# --------------------------------------------------------------------------
_DENIED = re.compile(r"\b(env|expandenv|getHostByName)\b")


def validate_template_string(value: str) -> Optional[str]:
    remaining = value
    while True:
        start = remaining.find("{{")
        if start == -1:
            return None
        remaining = remaining[start + 2:]
        end = remaining.find("}}")
        if end == -1:
            return None
        expr = remaining[:end].strip()
        if expr.startswith("-"):
            expr = expr[1:]
        if expr.endswith("-"):
            expr = expr[:-1]
        expr = expr.strip()
        if expr:
            m = _DENIED.search(expr)
            if m:
                for fn in ("env", "expandenv", "getHostByName"):
                    if re.search(r"\b%s\b" % fn, expr):
                        return "template expression uses disallowed function '%s'" % fn
                return "template expression uses disallowed function"
        remaining = remaining[end + 2:]


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def ensure_step_state_times(state: StepState, now: float) -> StepState:
    """step_state.go:25-33."""
    if state.phase != "" and state.started_at is None:
        state.started_at = now
    if is_terminal(state.phase) and state.finished_at is None:
        state.finished_at = now
    return state


def is_concurrency_queued(state: StepState) -> bool:
    """dag.go:2035-2051."""
    if state.phase != "Pending":
        return False
    return any(state.message.startswith(p) for p in QUEUED_PREFIXES)


def clear_concurrency_queued_steps(running: Dict[str, bool], states: Dict[str, StepState]) -> None:
    """dag.go:2020-2033."""
    if not running or not states:
        return
    for name in list(running.keys()):
        st = states.get(name)
        if st is None:
            continue
        if is_concurrency_queued(st):
            del running[name]


def all_story_steps(story: Story) -> List[Step]:
    """dag.go:3270-3280."""
    return list(story.steps) + list(story.compensations) + list(story.finally_)


def should_fail_fast(story: Story) -> bool:
    """dag.go:3504-3511."""
    if story.continue_on_step_failure is not None:
        return not story.continue_on_step_failure
    return True


def sanitize_step_identifier(name: str) -> str:
    """step_executor.go:1652-1670 (normalizeStepIdentifier dag.go:3548)."""
    return "".join(c if (c.isascii() and (c.isalnum() or c == "_")) else "_" for c in name)


# dag.go:3028-3030
STEP_NAME_REGEX = re.compile(
    r"""steps\.([a-zA-Z0-9_\-]+)\.|steps\s*\[\s*['"]([a-zA-Z0-9_\-]+)['"]\s*\]|\(index\s+\.steps\s+["']([a-zA-Z0-9_\-]+)["']\)"""
)


def _find_and_add_deps(expression: str, step_name: str, alias_to_real: Dict[str, str],
                       dependencies: Dict[str, Dict[str, bool]], dependents: Dict[str, Dict[str, bool]]) -> None:
    """dag.go:3223-3243."""
    for m in STEP_NAME_REGEX.finditer(expression):
        dep = m.group(3) or m.group(2) or m.group(1) or ""
        if dep == "":
            continue
        dep = alias_to_real.get(dep, dep)
        _add_dependency(step_name, dep, dependencies, dependents)


def _add_dependency(step_name: str, dep_name: str, dependencies, dependents) -> None:
    """dag.go:3262-3268."""
    dependencies[step_name][dep_name] = True
    dependents.setdefault(dep_name, {})[step_name] = True


def build_dependency_graphs(steps: List[Step]):
    """dag.go:3024-3073.  Returns (dependencies, dependents): name -> {name: True}."""
    dependencies: Dict[str, Dict[str, bool]] = {}
    dependents: Dict[str, Dict[str, bool]] = {}
    alias_to_real: Dict[str, str] = {}
    for s in steps:
        alias = sanitize_step_identifier(s.name)
        if alias != s.name:
            alias_to_real[alias] = s.name
    for step in steps:
        dependencies.setdefault(step.name, {})
        dependents.setdefault(step.name, {})
        for dep in step.needs:
            _add_dependency(step.name, dep, dependencies, dependents)
        if step.if_ is not None:
            _find_and_add_deps(step.if_, step.name, alias_to_real, dependencies, dependents)
        with_block = None
        if step.ref and step.with_ is not None:
            with_block = step.with_raw()
        elif step.type == "executeStory" and step.with_ is not None:
            with_block = step.with_raw()
        if with_block is not None:
            _find_and_add_deps(with_block, step.name, alias_to_real, dependencies, dependents)
    return dependencies, dependents


def validate_runtime_dependency_graph(steps: List[Step]) -> Optional[str]:
    """dag.go:3076-3146.  Returns None or the error text."""
    if not steps:
        return None
    dependencies, _ = build_dependency_graphs(steps)
    step_names = {s.name for s in steps}
    unknown = []
    for sname, deps in dependencies.items():
        for d in deps:
            if d not in step_names:
                unknown.append("%s->%s" % (sname, d))
    if unknown:
        unknown.sort()
        return "unknown step dependencies: " + ", ".join(unknown)
    indegree = {n: 0 for n in step_names}
    for sname, deps in dependencies.items():
        indegree[sname] = len(deps)
    ready = sorted(n for n, d in indegree.items() if d == 0)
    visited = 0
    while ready:
        current = ready.pop(0)
        visited += 1
        for sname, deps in dependencies.items():
            if not deps.get(current):
                continue
            indegree[sname] -= 1
            del deps[current]
            if indegree[sname] == 0:
                ready.append(sname)
                ready.sort()
    if visited == len(step_names):
        return None
    blocked = sorted(n for n, d in indegree.items() if d > 0)
    return "dependency cycle detected involving step(s): " + ", ".join(blocked)


def build_state_maps(steps: List[Step], states: Dict[str, StepState]):
    """dag.go:3358-3391 -> (completed, running, failed, allowedFailures)."""
    completed: Dict[str, bool] = {}
    running: Dict[str, bool] = {}
    failed: Dict[str, bool] = {}
    allowed_failures: Dict[str, bool] = {}
    allowed = {s.name for s in steps}
    allow_failure = {s.name: True for s in steps if s.allow_failure}
    for name, state in states.items():
        if name not in allowed:
            continue
        if state.phase in ("Succeeded", "Skipped"):
            completed[name] = True
        elif state.phase in ("Running", "Pending", "Paused"):
            running[name] = True
        elif is_terminal(state.phase):
            if allow_failure.get(name):
                completed[name] = True
                allowed_failures[name] = True
                continue
            failed[name] = True
    return completed, running, failed, allowed_failures


def steps_terminal(total: int, completed: Dict[str, bool], failed: Dict[str, bool]) -> bool:
    """dag.go:3282-3287."""
    if total == 0:
        return True
    return len(completed) + len(failed) == total


def mark_fail_fast_skipped(srun: StoryRun, story: Story, completed, running, now: float = 0.0) -> bool:
    """dag.go:3289-3312."""
    updated = False
    for step in story.steps:
        if completed.get(step.name) or running.get(step.name):
            continue
        state = srun.step_states.get(step.name, StepState()).copy()
        if is_terminal(state.phase):
            continue
        state.phase = "Skipped"
        state.message = "Skipped due to fail-fast policy"
        srun.step_states[step.name] = ensure_step_state_times(state, now)
        updated = True
    return updated


def mark_compensations_skipped(srun: StoryRun, story: Story, now: float = 0.0) -> bool:
    """dag.go:3314-3342."""
    if not story.compensations:
        return False
    completed, running, failed, _ = build_state_maps(story.compensations, srun.step_states)
    updated = False
    for step in story.compensations:
        if completed.get(step.name) or running.get(step.name) or failed.get(step.name):
            continue
        state = srun.step_states.get(step.name, StepState()).copy()
        if is_terminal(state.phase):
            continue
        state.phase = "Skipped"
        state.message = "Skipped because story succeeded"
        srun.step_states[step.name] = ensure_step_state_times(state, now)
        updated = True
    return updated


def collect_allowed_failures(steps: List[Step], states: Dict[str, StepState]) -> List[str]:
    """dag.go:3408-3430."""
    allow = {s.name for s in steps if s.allow_failure}
    out = []
    for name, st in states.items():
        if name not in allow:
            continue
        if st.phase in ("Succeeded", "Skipped"):
            continue
        if is_terminal(st.phase):
            out.append(name)
    return sorted(out)


def dependency_satisfied_for_realtime(story: Optional[Story], dep_state: StepState) -> bool:
    """dag.go:3457-3473."""
    if story is None or not story.realtime:
        return False
    if dep_state.phase == "":
        return False
    if is_terminal(dep_state.phase):
        return dep_state.phase == "Succeeded"
    return dep_state.phase in ("Pending", "Running", "Paused")


# --------------------------------------------------------------------------
# Go duration parsing (time.ParseDuration subset) for with.timeout etc.
# --------------------------------------------------------------------------
_DUR_UNITS = {"ns": 1e-9, "us": 1e-6, "µs": 1e-6, "ms": 1e-3, "s": 1.0, "m": 60.0, "h": 3600.0}
_DUR_RE = re.compile(r"([0-9]*\.?[0-9]+)(ns|us|µs|ms|s|m|h)")


def parse_positive_duration(raw: str) -> float:
    """steprun_controller.go:1364-1373."""
    pos, total = 0, 0.0
    s = raw
    if s == "" or s[0] in "+-" and len(s) == 1:
        raise ValueError("parse duration: invalid duration %r" % raw)
    sign = 1.0
    if s[0] in "+-":
        sign = -1.0 if s[0] == "-" else 1.0
        s = s[1:]
    if s == "0":
        total = 0.0
    else:
        while pos < len(s):
            m = _DUR_RE.match(s, pos)
            if not m:
                raise ValueError("parse duration: invalid duration %r" % raw)
            total += float(m.group(1)) * _DUR_UNITS[m.group(2)]
            pos = m.end()
    total *= sign
    if total <= 0:
        raise ValueError("duration must be positive")
    return total


def _normalize_on_timeout(raw: str) -> str:
    """dag.go:1643-1653."""
    if raw.strip() == "":
        return ""
    v = raw.strip().lower()
    if v in ("fail", "skip"):
        return v
    raise ValueError("unsupported value %r (expected fail or skip)" % raw)


def parse_gate_config(step: Step):
    """dag.go:1608-1641 -> (timeout|None, poll, onTimeout)."""
    if step.with_ is None:
        return None, 0.0, ""
    raw = step.with_ if isinstance(step.with_, dict) else {}
    timeout = None
    if raw.get("timeout"):
        timeout = parse_positive_duration(raw["timeout"])
    poll = parse_positive_duration(raw["pollInterval"]) if raw.get("pollInterval") else 0.0
    return timeout, poll, _normalize_on_timeout(raw.get("onTimeout", "") or "")


def parse_sleep_config(step: Step) -> float:
    """dag.go:1549-1567 -> duration (0 = none)."""
    if step.with_ is None:
        return 0.0
    raw = step.with_ if isinstance(step.with_, dict) else {}
    d = raw.get("duration", "") or ""
    if d.strip() == "":
        return 0.0
    return parse_positive_duration(d)


def parse_wait_config(step: Step):
    """dag.go:1569-1606 -> (until, timeout|None, poll, onTimeout)."""
    if step.with_ is None:
        raise ValueError("step '%s' of type 'wait' requires a 'with' block" % step.name)
    raw = step.with_
    until = raw.get("until", "") or ""
    if until.strip() == "":
        raise ValueError("step '%s' of type 'wait' requires 'with.until' to be set" % step.name)
    timeout = parse_positive_duration(raw["timeout"]) if raw.get("timeout") else None
    poll = parse_positive_duration(raw["pollInterval"]) if raw.get("pollInterval") else 0.0
    return until, timeout, poll, _normalize_on_timeout(raw.get("onTimeout", "") or "")


def apply_timeout_behavior(kind: str, on_timeout: str, state: StepState) -> None:
    """dag.go:1655-1668."""
    action = on_timeout.strip().lower() or "fail"
    if action == "skip":
        state.phase = "Skipped"
        state.message = "%s step timed out and was skipped." % kind
    else:
        state.phase = "Timeout"
        state.message = "%s step timed out." % kind


# --------------------------------------------------------------------------
# primitive syncs (stage G) and parallel join (stage H)
# --------------------------------------------------------------------------
def check_sync_gates(srun: StoryRun, story: Story, steps: List[Step], now: float = 0.0,
                     timers: Optional[StepTimers] = None) -> bool:
    """dag.go:1455-1547."""
    updated = False
    for step in steps:
        if step.type != "gate":
            continue
        current = srun.step_states.get(step.name)
        if current is None or is_terminal(current.phase):
            continue
        if current.phase not in ("Paused", "Running", "Pending"):
            continue
        try:
            timeout, _poll, on_timeout = parse_gate_config(step)
        except ValueError as e:
            nxt = ensure_step_state_times(StepState("Failed", str(e)), now)
            if nxt != current:
                srun.step_states[step.name] = nxt
                updated = True
            continue
        decision = "Pending"
        status = GateStatus()
        gs = srun.gates.get(step.name)
        if gs is not None:
            status = gs
            if gs.state != "":
                decision = gs.state
        nxt = current.copy()
        if decision == "Approved":
            nxt.phase = "Succeeded"
            nxt.message = status.message or "Gate approved."
        elif decision == "Rejected":
            nxt.phase = "Failed"
            nxt.message = status.message or "Gate rejected."
        else:
            nxt.phase = "Paused"
            if nxt.message == "":
                nxt.message = "Waiting for gate decision."
            if timeout is not None:
                if nxt.started_at is None:
                    nxt.started_at = now
                timeout_at = nxt.started_at + timeout
                if timers is not None:
                    timeout_at = timers.gate_timeout_at.setdefault(step.name, timeout_at)
                if not (now < timeout_at):
                    apply_timeout_behavior("gate", on_timeout, nxt)
        nxt = ensure_step_state_times(nxt, now)
        if nxt != current:
            srun.step_states[step.name] = nxt
            updated = True
    return updated


def check_sync_sleep_steps(srun: StoryRun, story: Story, steps: List[Step], now: float = 0.0,
                           timers: Optional[StepTimers] = None) -> bool:
    """dag.go:1217-1288."""
    updated = False
    for step in steps:
        if step.type != "sleep":
            continue
        current = srun.step_states.get(step.name)
        if current is None or is_terminal(current.phase):
            continue
        if current.phase not in ("Paused", "Running", "Pending"):
            continue
        try:
            duration = parse_sleep_config(step)
        except ValueError as e:
            nxt = ensure_step_state_times(StepState("Failed", str(e)), now)
            if nxt != current:
                srun.step_states[step.name] = nxt
                updated = True
            continue
        nxt = current.copy()
        if duration <= 0:
            nxt.phase = "Succeeded"
            nxt.message = SLEEP_COMPLETED_MSG
        else:
            if nxt.started_at is None:
                nxt.started_at = now
            sleep_until = nxt.started_at + duration
            if timers is not None:
                sleep_until = timers.sleep_until.setdefault(step.name, sleep_until)
            remaining = sleep_until - now
            if remaining <= 0:
                nxt.phase = "Succeeded"
                nxt.message = SLEEP_COMPLETED_MSG
            else:
                nxt.phase = "Paused"
                if nxt.message in ("", SLEEP_COMPLETED_MSG):
                    nxt.message = "Sleeping for %s." % duration
        nxt = ensure_step_state_times(nxt, now)
        if nxt != current:
            srun.step_states[step.name] = nxt
            updated = True
    return updated


def check_sync_wait_steps(srun: StoryRun, story: Story, steps: List[Step], evaluator: Optional[Evaluator],
                          vars_: Dict[str, Any], now: float = 0.0, timers: Optional[StepTimers] = None,
                          offloaded_policy: str = "fail") -> bool:
    """dag.go:1291-1452.  offloaded_policy: 'fail' | 'block' (controller/inject resolution is k8s I/O, out of scope)."""
    if evaluator is None:
        return False
    updated = False
    for step in steps:
        if step.type != "wait":
            continue
        current = srun.step_states.get(step.name)
        if current is None or is_terminal(current.phase):
            continue
        if current.phase not in ("Paused", "Running", "Pending"):
            continue
        try:
            until, timeout, _poll, on_timeout = parse_wait_config(step)
        except ValueError as e:
            nxt = ensure_step_state_times(StepState("Failed", str(e)), now)
            if nxt != current:
                srun.step_states[step.name] = nxt
                updated = True
            continue
        err = validate_template_string(until)
        if err is not None:
            nxt = ensure_step_state_times(StepState("Failed", err), now)
            if nxt != current:
                srun.step_states[step.name] = nxt
                updated = True
            continue
        result = False
        failed_state = None
        try:
            result = evaluator(step.name, until, vars_)
        except EvaluationBlocked:
            result = False
        except OffloadedDataUsage as e:
            if offloaded_policy == "block":
                result = False
            else:
                failed_state = StepState("Failed", str(e))
        except Exception:
            result = False
        if failed_state is not None:
            nxt = ensure_step_state_times(failed_state, now)
            if nxt != current:
                srun.step_states[step.name] = nxt
                updated = True
            continue
        nxt = current.copy()
        if result:
            nxt.phase = "Succeeded"
            nxt.message = "Wait condition satisfied."
        else:
            nxt.phase = "Paused"
            if nxt.message in ("", "Wait condition satisfied."):
                nxt.message = "Waiting for condition."
        waiting = not result
        if waiting and timeout is not None:
            if nxt.started_at is None:
                nxt.started_at = now
            timeout_at = nxt.started_at + timeout
            if timers is not None:
                timeout_at = timers.wait_timeout_at.setdefault(step.name, timeout_at)
            if not (now < timeout_at):
                apply_timeout_behavior("wait", on_timeout, nxt)
                waiting = False
        nxt = ensure_step_state_times(nxt, now)
        if nxt != current:
            srun.step_states[step.name] = nxt
            updated = True
    return updated


def parse_parallel_branches(step: Step) -> List[Step]:
    """dag.go:1202-1214.  with.steps entries are dicts {name, allowFailure?}."""
    if step.with_ is None:
        raise ValueError("parallel step '%s' missing 'with' configuration" % step.name)
    out = []
    for b in step.with_.get("steps", []):
        out.append(Step(name=b["name"], allow_failure=b.get("allowFailure"), ref=bool(b.get("ref", False))))
    return out


def check_sync_parallel_steps(srun: StoryRun, steps: List[Step], step_runs: Optional[List[StepRun]],
                              now: float = 0.0) -> bool:
    """dag.go:1112-1200."""
    if not steps or step_runs is None:
        return False
    if not srun.primitive_children:
        return False
    by_name = {sr.name: sr for sr in step_runs}
    updated = False
    for step in steps:
        if step.type != "parallel":
            continue
        state = srun.step_states.get(step.name)
        if state is None or is_terminal(state.phase):
            continue
        child_names = srun.primitive_children.get(step.name) or []
        if not child_names:
            continue
        allow_failure: Dict[str, bool] = {}
        if step.with_ is not None:
            try:
                for child in parse_parallel_branches(step):
                    if child.allow_failure:
                        allow_failure[child.name] = True
            except ValueError:
                pass
        all_done = True
        failed_branches: List[str] = []
        allowed_failures = 0
        for cname in child_names:
            child = by_name.get(cname)
            if child is None or child.phase == "" or not is_terminal(child.phase):
                all_done = False
                continue
            if child.phase in ("Succeeded", "Skipped"):
                continue
            if allow_failure.get(child.step_id):
                allowed_failures += 1
                continue
            failed_branches.append(child.step_id)
        if not all_done:
            continue
        nxt = state.copy()
        if failed_branches:
            nxt.phase = "Failed"
            nxt.message = "Parallel branches failed: " + ", ".join(failed_branches)
        elif allowed_failures > 0:
            nxt.phase = "Succeeded"
            nxt.message = "Parallel branches completed with allowed failures."
        else:
            nxt.phase = "Succeeded"
            nxt.message = "Parallel branches completed."
        srun.step_states[step.name] = ensure_step_state_times(nxt, now)
        updated = True
    return updated


# --------------------------------------------------------------------------
# stage D — findReadySteps
# --------------------------------------------------------------------------
@dataclass
class ReadyResult:
    ready: List[str]
    skipped: List[str]
    skip_reasons: Dict[str, str]
    unskipped: List[str]
    failed_now: List[str]        # steps whose state was set Failed by the `if` path (dag.go:2744, 2810)
    evaluated_if: List[str]      # steps whose `if` was evaluated (deps met) — the needs_cond set


def find_ready_steps(story: Optional[Story], steps: List[Step], step_states: Dict[str, StepState],
                     completed: Dict[str, bool], running: Dict[str, bool],
                     dependencies: Dict[str, Dict[str, bool]], vars_: Dict[str, Any],
                     dep_policy: DepPolicy, evaluator: Optional[Evaluator] = None,
                     dep_order: str = "failed_first", stale: Optional[Callable[[str, str], bool]] = None,
                     offloaded_policy: str = "fail", reevaluate_skipped: bool = False,
                     now: float = 0.0) -> ReadyResult:
    """dag.go:2631-2848.

    dep_order pins Go's random map iteration over deps (dag.go:2714):
      'failed_first' -> a failed dep is met before any unmet dep  => skip_max
      'unmet_first'  -> an unmet dep is met before any failed dep => skip_min
      'insertion'    -> dict order.
    `stale(step, 'if'|'with')` stands in for outputRefsMaybeStale/withRefsMaybeStale (dag.go:3151-3202),
    which look into JSON outputs (host work).  reevaluate_skipped enables the :2654-2705 branch.
    """
    ready: List[str] = []
    skipped: List[str] = []
    skip_reasons: Dict[str, str] = {}
    unskipped: List[str] = []
    failed_now: List[str] = []
    evaluated_if: List[str] = []
    is_realtime = story is not None and story.realtime
    stale = stale or (lambda _s, _k: False)

    for step in steps:
        if completed.get(step.name) or running.get(step.name):
            continue
        st = step_states.get(step.name)
        if st is not None and is_terminal(st.phase):
            if reevaluate_skipped and st.phase == "Skipped" and step.if_:
                if not stale(step.name, "if"):
                    try:
                        result = evaluator(step.name, step.if_, vars_) if evaluator else False
                    except Exception:
                        continue
                    if result:
                        del step_states[step.name]
                        completed[step.name] = False
                        unskipped.append(step.name)
                    else:
                        continue
                else:
                    continue
            else:
                continue

        deps = list(dependencies.get(step.name, {}).keys())

        def _klass(dep: str) -> int:
            # 0 satisfied, 1 failed-dep, 2 unmet — same clause order as dag.go:2715-2732
            if completed.get(dep):
                return 0
            ds = step_states.get(dep, StepState())
            if dependency_satisfied_for_realtime(story, ds):
                return 0
            if dep_policy.allow_failed_dependencies and is_terminal(ds.phase):
                return 0
            if dep_policy.skip_on_failed_dependency and is_terminal(ds.phase) and ds.phase not in ("Succeeded", "Skipped"):
                return 1
            return 2

        if dep_order == "failed_first":
            deps.sort(key=lambda d: (0 if _klass(d) == 1 else 1))
        elif dep_order == "unmet_first":
            deps.sort(key=lambda d: (0 if _klass(d) == 2 else 1))

        all_deps_met = True
        failed_dep = ""
        for dep in deps:
            k = _klass(dep)
            if k == 0:
                continue
            if k == 1:
                failed_dep = dep
                break
            all_deps_met = False
            break

        if failed_dep != "":
            skipped.append(step.name)
            skip_reasons[step.name] = "Skipped due to failed dependency: %s" % failed_dep
            continue
        if not all_deps_met:
            continue

        if step.if_ and not is_realtime:
            evaluated_if.append(step.name)
            safety = validate_template_string(step.if_)
            if safety is not None:
                step_states[step.name] = ensure_step_state_times(StepState("Failed", safety), now)
                failed_now.append(step.name)
                continue
            result = False
            try:
                result = evaluator(step.name, step.if_, vars_) if evaluator else False
            except EvaluationBlocked:
                continue
            except OffloadedDataUsage as e:
                if offloaded_policy in ("block", "controller", "inject"):
                    continue
                step_states[step.name] = ensure_step_state_times(StepState("Failed", str(e)), now)
                failed_now.append(step.name)
                continue
            except Exception:
                continue
            if not result:
                if stale(step.name, "if"):
                    continue
                skipped.append(step.name)
                skip_reasons[step.name] = "Skipped due to 'if' condition"
                continue
        if stale(step.name, "with"):
            continue
        ready.append(step.name)
    return ReadyResult(ready, skipped, skip_reasons, unskipped, failed_now, evaluated_if)


# --------------------------------------------------------------------------
# launch effects (the host side of findAndLaunchReadySteps)
# --------------------------------------------------------------------------
def apply_launch_effects(srun: StoryRun, story: Story, res: ReadyResult, by_name: Dict[str, Step],
                         now: float = 0.0, device_contract: bool = False) -> List[Tuple[str, str]]:
    """dag.go:1735-1775 + step_executor.go:132-185, 740-811, 1081-1106.

    Limiters (enforceStoryConcurrency/SchedulingLimits, dag.go:1713-1728) need
    cluster-wide LISTs and are a 'next' row; this models unlimited slots.
    device_contract=True reproduces the packed fixpoint contract: a ready `stop` step is
    NOT executed on the device (its phase comes from with.phase, host data) — the run is handed to the host.
    Returns the (parallel step, branch) expansion list in creation order."""
    expansion: List[Tuple[str, str]] = []
    for name in res.skipped:
        cur = srun.step_states.get(name)
        if cur is None or cur.phase != "Skipped":
            msg = res.skip_reasons.get(name) or "Skipped due to 'if' condition"
            srun.step_states[name] = ensure_step_state_times(StepState("Skipped", msg), now)
    for name in res.ready:
        step = by_name[name]
        if step.ref:
            pass  # executeEngramStep: creates a StepRun (k8s I/O); state set below
        elif step.type == "executeStory":
            pass
        elif step.type == "parallel":
            children = []
            for b in parse_parallel_branches(step):
                children.append("%s-%s" % (step.name, b.name))
                expansion.append((step.name, b.name))
            srun.primitive_children[step.name] = children
            _mark_step_state(srun, name, "Running", "Parallel block expanded", now)
        elif step.type == "condition":
            _mark_step_state(srun, name, "Succeeded", "Primitive evaluated and outputs are available.", now)
        elif step.type == "sleep":
            _mark_step_state(srun, name, "Paused", "Sleeping.", now)
        elif step.type == "gate":
            _mark_step_state(srun, name, "Paused", "Waiting for gate decision.", now)
        elif step.type == "wait":
            _mark_step_state(srun, name, "Paused", "Waiting for condition.", now)
        elif step.type == "stop" and device_contract:
            continue
        elif step.type == "stop":
            w = step.with_ or {}
            phase = w.get("phase") or "Succeeded"
            msg = w.get("message") or "Story execution stopped by step '%s' with phase '%s'" % (step.name, phase)
            srun.phase = phase
            _mark_step_state(srun, name, phase, msg, now)
        else:
            raise ValueError("step '%s' has an unsupported type '%s' or is missing a 'ref'" % (step.name, step.type))
        st = srun.step_states.get(name)
        if st is None or st.phase == "" or is_concurrency_queued(st):
            srun.step_states[name] = ensure_step_state_times(StepState("Running"), now)
    return expansion


def _mark_step_state(srun: StoryRun, name: str, phase: str, message: str, now: float) -> None:
    """step_executor.go:1236-1247."""
    st = srun.step_states.get(name, StepState()).copy()
    st.phase = phase
    st.message = message
    srun.step_states[name] = ensure_step_state_times(st, now)


# --------------------------------------------------------------------------
# one iteration of runDagIterations (stages G,H,I,D) and the fixpoint (J)
# --------------------------------------------------------------------------
@dataclass
class IterationResult:
    group: str                     # "main" | "compensation" | "finally" | "finalize"
    main_done: bool
    main_failed: bool
    comp_done: bool
    final_done: bool
    comp_failed: bool
    final_failed: bool
    ready: ReadyResult
    ready_min: Optional[ReadyResult] = None   # skip_min variant


def run_dag_iteration(srun: StoryRun, story: Story, step_runs: Optional[List[StepRun]] = None,
                      evaluator: Optional[Evaluator] = None, vars_: Optional[Dict[str, Any]] = None,
                      now: float = 0.0, timers: Optional[StepTimers] = None,
                      stale: Optional[Callable[[str, str], bool]] = None,
                      host_group: Optional[str] = None, dep_order: str = "failed_first",
                      offloaded_policy: str = "fail") -> IterationResult:
    """One pass of the loop body dag.go:393-540 up to and including findReadySteps
    (no launch effects).  host_group != None reproduces contract tier K1 (BF_RF_HOST_GROUP): the caller
    supplies the group and stage I (fail-fast / compensation marking, group selection) is skipped."""
    vars_ = vars_ if vars_ is not None else {"inputs": {}, "steps": {}}
    all_steps = all_story_steps(story)
    check_sync_gates(srun, story, all_steps, now, timers)                      # dag.go:409
    check_sync_sleep_steps(srun, story, all_steps, now, timers)                # :412
    check_sync_wait_steps(srun, story, all_steps, evaluator, vars_, now, timers, offloaded_policy)  # :415
    check_sync_parallel_steps(srun, all_steps, step_runs, now)                 # :418
    if host_group is None:
        main_completed, main_running, main_failed, _ = build_state_maps(story.steps, srun.step_states)  # :422
        clear_concurrency_queued_steps(main_running, srun.step_states)
        if should_fail_fast(story) and len(main_failed) > 0:                        # :424
            mark_fail_fast_skipped(srun, story, main_completed, main_running, now)
            main_completed, main_running, main_failed, _ = build_state_maps(story.steps, srun.step_states)
            clear_concurrency_queued_steps(main_running, srun.step_states)
        main_done = steps_terminal(len(story.steps), main_completed, main_failed)  # :431

        if not main_done and story.realtime and srun.topology_terminated:          # :436-464
            main_done = True
            for step in story.steps:
                ss = srun.step_states.get(step.name)
                if ss is not None and not is_terminal(ss.phase):
                    ss = ss.copy()
                    ss.phase = "Failed"
                    ss.message = "realtime topology terminated"
                    srun.step_states[step.name] = ss
                    main_failed[step.name] = True
            _, main_running, main_failed, _ = build_state_maps(story.steps, srun.step_states)
            clear_concurrency_queued_steps(main_running, srun.step_states)

        if main_done and len(main_failed) == 0 and len(story.compensations) > 0:   # :466
            mark_compensations_skipped(srun, story, now)

        comp_completed, _, comp_failed, _ = build_state_maps(story.compensations, srun.step_states)  # :472
        final_completed, _, final_failed, _ = build_state_maps(story.finally_, srun.step_states)
        comp_done = steps_terminal(len(story.compensations), comp_completed, comp_failed)
        final_done = steps_terminal(len(story.finally_), final_completed, final_failed)
        srun.allowed_failures = collect_allowed_failures(all_steps, srun.step_states)   # :478

        if not main_done:                                                           # :482-495
            group = "main"
        elif len(main_failed) > 0 and len(story.compensations) > 0 and not comp_done:
            group = "compensation"
        elif len(story.finally_) > 0 and not final_done:
            group = "finally"
        else:
            group = "finalize"
        flags = dict(main_done=main_done, main_failed=len(main_failed) > 0, comp_done=comp_done,
                     final_done=final_done, comp_failed=len(comp_failed) > 0, final_failed=len(final_failed) > 0)
    else:
        group = host_group
        flags = dict(main_done=False, main_failed=False, comp_done=False, final_done=False,
                     comp_failed=False, final_failed=False)

    empty = ReadyResult([], [], {}, [], [], [])
    if group == "finalize":
        return IterationResult(group, ready=empty, ready_min=empty, **flags)

    completed, running, _, _ = build_state_maps(all_steps, srun.step_states)        # :497
    clear_concurrency_queued_steps(running, srun.step_states)
    dep_policy = DepPolicy(allow_failed_dependencies=(group != "main"),             # :499-502
                           skip_on_failed_dependency=(group == "main" and not should_fail_fast(story)))
    step_list = {"compensation": story.compensations, "finally": story.finally_}.get(group, story.steps)
    dependencies, _ = build_dependency_graphs(step_list)                            # :1700
    # skip_min variant first, on copies (find_ready_steps may mutate step_states on `if` failures)
    states_min = {k: v.copy() for k, v in srun.step_states.items()}
    res_min = find_ready_steps(story, step_list, states_min, dict(completed), dict(running), dependencies,
                               vars_, dep_policy, evaluator, "unmet_first", stale, offloaded_policy, False, now)
    res = find_ready_steps(story, step_list, srun.step_states, completed, running, dependencies, vars_,
                           dep_policy, evaluator, dep_order, stale, offloaded_policy, False, now)
    return IterationResult(group, ready=res, ready_min=res_min, **flags)


def run_dag_iterations(srun: StoryRun, story: Story, step_runs: Optional[List[StepRun]] = None,
                       evaluator: Optional[Evaluator] = None, vars_: Optional[Dict[str, Any]] = None,
                       now: float = 0.0, timers: Optional[StepTimers] = None,
                       stale: Optional[Callable[[str, str], bool]] = None,
                       max_iterations: Optional[int] = None, device_contract: bool = False):
    """dag.go:381-542 (fixpoint J) with launch effects and unlimited concurrency slots.
    device_contract: a ready `stop` step ends the loop after its iteration (see apply_launch_effects).
    Returns (iterations_run, launched, skipped, expansion, final IterationResult)."""
    all_steps = all_story_steps(story)
    err = validate_runtime_dependency_graph(all_steps)
    if err is not None:
        raise ValueError("Invalid story dependency graph: " + err)
    by_name = {s.name: s for s in all_steps}
    launched: List[str] = []
    skipped: List[str] = []
    expansion: List[Tuple[str, str]] = []
    iters = 0
    last = None
    cap = len(all_steps) + 1 if max_iterations is None else max_iterations
    for _ in range(cap):
        last = run_dag_iteration(srun, story, step_runs, evaluator, vars_, now, timers, stale)
        iters += 1
        if last.group == "finalize":
            break
        expansion += apply_launch_effects(srun, story, last.ready, by_name, now, device_contract)
        launched += last.ready.ready
        skipped += last.ready.skipped
        if len(last.ready.ready) == 0 and len(last.ready.skipped) == 0:
            break
        if device_contract and any(by_name[n].type == "stop" and not by_name[n].ref for n in last.ready.ready):
            break
    return iters, launched, skipped, expansion, last


# --------------------------------------------------------------------------
# internal/controller/runs/storyrun_controller.go:535-577 — redrive closure
# --------------------------------------------------------------------------
def find_step_group(story: Story, step_name: str) -> Optional[List[Step]]:
    """findStepGroup, storyrun_controller.go:560-577."""
    for group in (story.steps, story.compensations, story.finally_):
        if any(s.name == step_name for s in group):
            return group
    return None


def resolve_redrive_from_step_set(story: Story, step_name: str) -> Dict[str, bool]:
    """resolveRedriveFromStepSet, storyrun_controller.go:535-558: BFS over the dependents map of the step's own group."""
    steps = find_step_group(story, step_name)
    if steps is None:
        raise KeyError("step %r not found in story" % step_name)
    _, dependents = build_dependency_graphs(steps)
    selected = {step_name: True}
    queue = [step_name]
    while queue:
        current = queue.pop(0)
        for dep in dependents.get(current, {}):
            if selected.get(dep):
                continue
            selected[dep] = True
            queue.append(dep)
    return selected
