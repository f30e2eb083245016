"""Known-answer tests that PIN the oracle.

Every test re-expresses one of the reference's own tests for the DAG frontier path
(/root/reference/internal/controller/runs/dag_test.go, cited per test) against
oracle/pyoracle.py, and — where the case fits the packed contract — against the packed
C oracle (oracle/packed_ref.c) through the test-side packer.  Config 1 of BASELINE.json
(A->B->C) has no dedicated reference test; its snapshots are pinned here too.
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from oracle import packed as PK
from bobrapet_b200 import _abi as A
from bobrapet_b200.records import unpack_result
from tests import packing as P


# Backend of _packed_pass: "oracle" (CPU, pins the oracle) or "cuda" (marked gpu): the SAME transcribed dag_test.go
# vectors go straight through bf_eval on the GPU and the reference's expected answers are asserted on the CUDA output
# (the records are also compared byte for byte with the packed oracle).
_BACKEND = {"name": "oracle", "frontier": None, "calls": 0}


def _uses_packed_pass():
    """names of this module's functions that reach _packed_pass (directly or through a helper)"""
    import ast
    import inspect
    import sys
    tree = ast.parse(inspect.getsource(sys.modules[__name__]))
    calls = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            calls[node.name] = {n.func.id for n in ast.walk(node) if isinstance(n, ast.Call) and isinstance(n.func, ast.Name)}
    reach = {"_packed_pass"}
    changed = True
    while changed:
        changed = False
        for fn, cs in calls.items():
            if fn not in reach and cs & reach:
                reach.add(fn)
                changed = True
    return reach


_REACH = None


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)], autouse=True)
def backend(request):
    global _REACH
    if request.param == "cuda":
        if _REACH is None:
            _REACH = _uses_packed_pass()
        if request.function.__name__ not in _REACH:
            pytest.skip("no packed vector in this test (oracle-only KAT)")
        from bobrapet_b200 import Frontier
        if _BACKEND["frontier"] is None:
            _BACKEND["frontier"] = Frontier(0)
    _BACKEND["name"], _BACKEND["calls"] = request.param, 0
    yield request.param
    if request.param == "cuda":
        assert _BACKEND["calls"] > 0, "cuda variant ran without a bf_eval call"
    _BACKEND["name"] = "oracle"


def _packed_pass(story, srun, step_runs=None, evaluator=None, vars_=None, now=0.0, timers=None,
                 host_group=None, flags=0):
    ps = P.pack_story(story)
    ts = P.topology_set([ps])
    L, state = P.pack_runs([story], [ps], [srun], [0], [0], [step_runs], evaluator, vars_, now,
                           [timers], None, [host_group])
    if _BACKEND["name"] == "cuda":
        fr = _BACKEND["frontier"]
        slots = fr.put_topologies(ts)
        try:
            state[:, 0:4] = np.ascontiguousarray(slots[:1], dtype="<u4").view(np.uint8)
            res, counts = fr.eval(L, state, flags=flags | A.EVAL_VALIDATE)
            want, wcounts = PK.evaluate(PK.PackedTopologies(ts, slots), L, state, flags)
            assert np.array_equal(res, want) and counts == wcounts, "bf_eval differs from oracle/packed_ref.c on a reference KAT"
        finally:
            fr.drop_topology(int(slots[0]))
        _BACKEND["calls"] += 1
    else:
        res, counts = PK.evaluate(PK.PackedTopologies(ts), L, state, flags)
    out = unpack_result(L, res, ps.S)
    names = ps.names
    pick = lambda key: [names[i] for i in np.nonzero(out[key][0])[0]]
    phases = {names[i]: A.PHASE_NAMES[c] for i, c in enumerate(out["phase_out"][0])}
    return {"ready": pick("ready"), "skip": pick("skip"), "fail": pick("fail"), "needs_cond": pick("needs_cond"),
            "skip_dep": pick("skip_dep"), "phase": phases, "summary": int(out["summary"][0]), "counts": counts,
            "n_expansion": int(out["n_expansion"][0])}


# ---- dag_test.go:78 TestBuildStateMapsAllowsFailures
def test_build_state_maps_allows_failures():
    steps = [O.Step("allowed", allow_failure=True), O.Step("blocked")]
    states = {"allowed": O.StepState("Failed"), "blocked": O.StepState("Failed")}
    completed, running, failed, allowed = O.build_state_maps(steps, states)
    assert running == {}
    assert completed.get("allowed") and not failed.get("allowed")
    assert failed.get("blocked")
    assert allowed.get("allowed")


# ---- dag_test.go:1172 TestBuildStateMapsIgnoresNonStorySteps
def test_build_state_maps_ignores_non_story_steps():
    steps = [O.Step("fetch-feeds"), O.Step("generate-digest")]
    states = {"fetch-feeds": O.StepState("Succeeded"), "generate-digest": O.StepState("Running"),
              "fetch-feed-batch": O.StepState("Succeeded")}
    completed, running, failed, _ = O.build_state_maps(steps, states)
    assert failed == {}
    assert completed == {"fetch-feeds": True}
    assert running == {"generate-digest": True}


# ---- dag_test.go:1200 TestBuildStateMapsTreatsPausedAsRunning
def test_build_state_maps_paused_is_running():
    completed, running, failed, _ = O.build_state_maps([O.Step("gate")], {"gate": O.StepState("Paused")})
    assert completed == {} and failed == {} and running == {"gate": True}


# ---- dag_test.go:107 TestMarkCompensationsSkipped
def test_mark_compensations_skipped():
    story = O.Story(compensations=[O.Step("rollback-a"), O.Step("rollback-b")])
    srun = O.StoryRun(step_states={"rollback-a": O.StepState("Succeeded")})
    assert O.mark_compensations_skipped(srun, story)
    assert srun.step_states["rollback-a"].phase == "Succeeded"
    assert srun.step_states["rollback-b"].phase == "Skipped"


# ---- dag_test.go:842 TestFindReadyStepsSkipsFailedDependencies
def test_find_ready_steps_skips_failed_dependencies():
    story = O.Story(steps=[O.Step("fetch-feed"), O.Step("extract-items", needs=["fetch-feed"])],
                    continue_on_step_failure=True)
    srun = O.StoryRun(step_states={"fetch-feed": O.StepState("Failed")})
    completed, running, _, _ = O.build_state_maps(story.steps, srun.step_states)
    deps, _ = O.build_dependency_graphs(story.steps)
    res = O.find_ready_steps(story, story.steps, srun.step_states, completed, running, deps, {},
                             O.DepPolicy(skip_on_failed_dependency=True))
    assert res.ready == []
    assert res.skipped == ["extract-items"]
    assert "failed dependency" in res.skip_reasons["extract-items"]
    # the same vector through the packed contract (main group, !failFast => skipOnFailedDependency)
    got = _packed_pass(story, O.StoryRun(step_states={"fetch-feed": O.StepState("Failed")}))
    assert got["ready"] == [] and got["skip"] == ["extract-items"] and got["skip_dep"] == ["extract-items"]


# ---- dag_test.go:690 TestFindAndLaunchReadyStepsInitializesStepStates  (`if: "false"` condition step)
def test_if_false_condition_step_is_skipped():
    story = O.Story(steps=[O.Step("skip-me", type="condition", if_="false")])
    srun = O.StoryRun()
    it = O.run_dag_iteration(srun, story, evaluator=O.literal_evaluator)
    assert it.ready.ready == [] and it.ready.skipped == ["skip-me"]
    O.apply_launch_effects(srun, story, it.ready, {"skip-me": story.steps[0]})
    assert srun.step_states["skip-me"].phase == "Skipped"
    got = _packed_pass(story, O.StoryRun(), evaluator=O.literal_evaluator, flags=A.EVAL_FIXPOINT)
    assert got["ready"] == [] and got["skip"] == ["skip-me"] and got["phase"]["skip-me"] == "Skipped"


# ---- dag_test.go:886 / :1028 / :1084 TestCheckSyncParallelSteps*
def _parallel_fixture(b_phase, allow_b=False):
    branches = [{"name": "branch-a"}, {"name": "branch-b", **({"allowFailure": True} if allow_b else {})}]
    step = O.Step("parallel", type="parallel", with_={"steps": branches})
    srun = O.StoryRun(primitive_children={"parallel": ["parallel-branch-a", "parallel-branch-b"]},
                      step_states={"parallel": O.StepState("Running")})
    srs = [O.StepRun("parallel-branch-a", "branch-a", "Succeeded"), O.StepRun("parallel-branch-b", "branch-b", b_phase)]
    return step, srun, srs


@pytest.mark.parametrize("b_phase,allow_b,expect", [("Succeeded", False, "Succeeded"), ("Failed", False, "Failed"),
                                                    ("Failed", True, "Succeeded")])
def test_check_sync_parallel_steps(b_phase, allow_b, expect):
    step, srun, srs = _parallel_fixture(b_phase, allow_b)
    assert O.check_sync_parallel_steps(srun, [step], srs)
    assert srun.step_states["parallel"].phase == expect
    step, srun, srs = _parallel_fixture(b_phase, allow_b)
    got = _packed_pass(O.Story(steps=[step]), srun, step_runs=srs)
    assert got["phase"]["parallel"] == expect


def test_check_sync_parallel_waits_for_unfinished_child():
    step, srun, srs = _parallel_fixture("Running")
    assert not O.check_sync_parallel_steps(srun, [step], srs)
    assert srun.step_states["parallel"].phase == "Running"
    got = _packed_pass(O.Story(steps=[step]), srun, step_runs=srs)
    assert got["phase"]["parallel"] == "Running"


# ---- dag_test.go:1141 TestDependencySatisfiedForRealtime
@pytest.mark.parametrize("phase,expect", [("Pending", True), ("Running", True), ("Paused", True), ("Succeeded", True),
                                          ("Failed", False), ("Canceled", False), ("", False)])
def test_dependency_satisfied_for_realtime(phase, expect):
    rt = O.Story(realtime=True)
    assert O.dependency_satisfied_for_realtime(rt, O.StepState(phase)) is expect
    # batch stories never short-circuit
    assert O.dependency_satisfied_for_realtime(O.Story(), O.StepState("Running")) is False
    # and through the packed contract: B needs A; A in `phase`
    story = O.Story(steps=[O.Step("A", ref=True), O.Step("B", ref=True, needs=["A"])], realtime=True)
    srun = O.StoryRun(step_states=({"A": O.StepState(phase)} if phase else {}))
    got = _packed_pass(story, srun)
    assert ("B" in got["ready"]) is expect


# ---- dag_test.go:1223 / :1253 TestCheckSyncGatesApproved / Rejected
@pytest.mark.parametrize("decision,expect,msg", [("Approved", "Succeeded", "ok"), ("Rejected", "Failed", "no")])
def test_check_sync_gates(decision, expect, msg):
    story = O.Story(steps=[O.Step("approve", type="gate")])
    srun = O.StoryRun(step_states={"approve": O.StepState("Paused")}, gates={"approve": O.GateStatus(decision, msg)})
    assert O.check_sync_gates(srun, story, story.steps)
    assert srun.step_states["approve"].phase == expect
    assert srun.step_states["approve"].message == msg
    srun = O.StoryRun(step_states={"approve": O.StepState("Paused")}, gates={"approve": O.GateStatus(decision, msg)})
    assert _packed_pass(story, srun)["phase"]["approve"] == expect


# ---- dag_test.go:1416 / :1447 / :1480 TestCheckSyncSleepSteps{Completes,Initializes,UsesStoredDeadline}
def test_check_sync_sleep_completes():
    story = O.Story(steps=[O.Step("nap", type="sleep", with_={"duration": "1s"})])
    srun = O.StoryRun(step_states={"nap": O.StepState("Paused", started_at=100.0)})
    assert O.check_sync_sleep_steps(srun, story, story.steps, now=102.0)
    assert srun.step_states["nap"].phase == "Succeeded"
    srun = O.StoryRun(step_states={"nap": O.StepState("Paused", started_at=100.0)})
    assert _packed_pass(story, srun, now=102.0)["phase"]["nap"] == "Succeeded"


def test_check_sync_sleep_initializes():
    story = O.Story(steps=[O.Step("nap", type="sleep", with_={"duration": "10s"})])
    srun = O.StoryRun(step_states={"nap": O.StepState("Paused")})
    assert O.check_sync_sleep_steps(srun, story, story.steps, now=50.0)
    st = srun.step_states["nap"]
    assert st.phase == "Paused" and st.started_at == 50.0
    srun = O.StoryRun(step_states={"nap": O.StepState("Paused")})
    assert _packed_pass(story, srun, now=50.0)["phase"]["nap"] == "Paused"


def test_check_sync_sleep_uses_stored_deadline():
    story = O.Story(steps=[O.Step("nap", type="sleep", with_={"duration": "10s"})])
    timers = O.StepTimers(sleep_until={"nap": 101.0})
    srun = O.StoryRun(step_states={"nap": O.StepState("Paused", started_at=100.0)})
    assert O.check_sync_sleep_steps(srun, story, story.steps, now=102.0, timers=timers)
    assert srun.step_states["nap"].phase == "Succeeded"   # stored deadline wins over startedAt+duration
    srun = O.StoryRun(step_states={"nap": O.StepState("Paused", started_at=100.0)})
    assert _packed_pass(story, srun, now=102.0, timers=timers)["phase"]["nap"] == "Succeeded"


# ---- dag_test.go:1283 / :1322 / :1367 TestCheckSyncWaitSteps{Satisfied,TimeoutSkip,UsesStoredTimeout}
def test_check_sync_wait_satisfied():
    story = O.Story(steps=[O.Step("wait", type="wait", with_={"until": "{{ inputs.ready }}"})])
    vars_ = {"inputs": {"ready": True}, "steps": {}}
    srun = O.StoryRun(step_states={"wait": O.StepState("Paused")})
    assert O.check_sync_wait_steps(srun, story, story.steps, O.literal_evaluator, vars_)
    assert srun.step_states["wait"].phase == "Succeeded"
    srun = O.StoryRun(step_states={"wait": O.StepState("Paused")})
    assert _packed_pass(story, srun, evaluator=O.literal_evaluator, vars_=vars_)["phase"]["wait"] == "Succeeded"


def test_check_sync_wait_timeout_skip():
    story = O.Story(steps=[O.Step("wait", type="wait", with_={"until": "{{ inputs.ready }}", "timeout": "1s", "onTimeout": "skip"})])
    vars_ = {"inputs": {"ready": False}, "steps": {}}
    srun = O.StoryRun(step_states={"wait": O.StepState("Paused", started_at=100.0)})
    assert O.check_sync_wait_steps(srun, story, story.steps, O.literal_evaluator, vars_, now=102.0)
    assert srun.step_states["wait"].phase == "Skipped"
    srun = O.StoryRun(step_states={"wait": O.StepState("Paused", started_at=100.0)})
    assert _packed_pass(story, srun, evaluator=O.literal_evaluator, vars_=vars_, now=102.0)["phase"]["wait"] == "Skipped"


def test_check_sync_wait_uses_stored_timeout():
    story = O.Story(steps=[O.Step("wait", type="wait", with_={"until": "{{ inputs.ready }}", "timeout": "10s"})])
    vars_ = {"inputs": {"ready": False}, "steps": {}}
    timers = O.StepTimers(wait_timeout_at={"wait": 101.0})
    srun = O.StoryRun(step_states={"wait": O.StepState("Paused", started_at=100.0)})
    assert O.check_sync_wait_steps(srun, story, story.steps, O.literal_evaluator, vars_, now=102.0, timers=timers)
    assert srun.step_states["wait"].phase == "Timeout"
    srun = O.StoryRun(step_states={"wait": O.StepState("Paused", started_at=100.0)})
    got = _packed_pass(story, srun, evaluator=O.literal_evaluator, vars_=vars_, now=102.0, timers=timers)
    assert got["phase"]["wait"] == "Timeout"


# ---- dag_test.go:206 / :321 TestDAGReconcileFailsOnUnknownDependency / OnDependencyCycle
def test_validate_unknown_dependency():
    err = O.validate_runtime_dependency_graph([O.Step("a", needs=["ghost"])])
    assert err is not None and "unknown step dependencies: a->ghost" in err


def test_validate_dependency_cycle_self_loop():
    err = O.validate_runtime_dependency_graph([O.Step("loop", needs=["loop"])])
    assert err is not None and "dependency cycle detected involving step(s): loop" in err


def test_validate_two_cycle_and_ok_chain():
    assert O.validate_runtime_dependency_graph([O.Step("a", needs=["b"]), O.Step("b", needs=["a"])]) is not None
    assert O.validate_runtime_dependency_graph([O.Step("a"), O.Step("b", needs=["a"]), O.Step("c", needs=["b"])]) is None


# ---- dag_test.go:1603 / :1707 / :1798 realtime topology terminated
def test_realtime_topology_terminated_triggers_compensation():
    story = O.Story(steps=[O.Step("stream", ref=True)], compensations=[O.Step("undo", ref=True)], realtime=True)
    srun = O.StoryRun(step_states={"stream": O.StepState("Running")}, topology_terminated=True)
    it = O.run_dag_iteration(srun, story)
    assert srun.step_states["stream"].phase == "Failed"
    assert it.group == "compensation" and it.ready.ready == ["undo"]
    srun = O.StoryRun(step_states={"stream": O.StepState("Running")}, topology_terminated=True)
    got = _packed_pass(story, srun)
    assert got["phase"]["stream"] == "Failed" and got["ready"] == ["undo"]
    assert got["summary"] & A.SUM_GROUP_MASK == A.GROUP_COMPENSATION


def test_realtime_topology_terminated_triggers_finally():
    story = O.Story(steps=[O.Step("stream", ref=True)], finally_=[O.Step("cleanup", ref=True)], realtime=True)
    srun = O.StoryRun(step_states={"stream": O.StepState("Running")}, topology_terminated=True)
    it = O.run_dag_iteration(srun, story)
    assert it.group == "finally" and it.ready.ready == ["cleanup"]
    srun = O.StoryRun(step_states={"stream": O.StepState("Running")}, topology_terminated=True)
    got = _packed_pass(story, srun)
    assert got["ready"] == ["cleanup"] and got["summary"] & A.SUM_GROUP_MASK == A.GROUP_FINALLY


def test_realtime_not_degraded_skips_termination():
    story = O.Story(steps=[O.Step("stream", ref=True)], finally_=[O.Step("cleanup", ref=True)], realtime=True)
    srun = O.StoryRun(step_states={"stream": O.StepState("Running")}, topology_terminated=False)
    it = O.run_dag_iteration(srun, story)
    assert it.group == "main" and srun.step_states["stream"].phase == "Running" and it.ready.ready == []
    got = _packed_pass(story, O.StoryRun(step_states={"stream": O.StepState("Running")}))
    assert got["ready"] == [] and got["phase"]["stream"] == "Running"


# ---- webhook KAT: a->b->c needs chain (story_webhook_test.go:975-977) as a runtime graph
def test_needs_chain_graph_shape():
    deps, dependents = O.build_dependency_graphs([O.Step("a"), O.Step("b", needs=["a"]), O.Step("c", needs=["b"])])
    assert deps == {"a": {}, "b": {"a": True}, "c": {"b": True}}
    assert dependents["a"] == {"b": True} and dependents["b"] == {"c": True}


def test_implicit_dependencies_from_templates():
    steps = [O.Step("fetch-feed", ref=True),
             O.Step("use-dot", ref=True, with_={"x": "{{ steps.fetch_feed.output.body }}"}),
             O.Step("use-index", ref=True, if_='{{ (index .steps "fetch-feed").output.ok }}'),
             O.Step("use-bracket", type="executeStory", with_={"y": "{{ steps['fetch-feed'].output }}"}),
             O.Step("ignored", type="condition", with_={"z": "{{ steps.fetch_feed.output }}"})]
    deps, _ = O.build_dependency_graphs(steps)
    assert deps["use-dot"] == {"fetch-feed": True}        # alias fetch_feed -> fetch-feed (dag.go:3033-3039)
    assert deps["use-index"] == {"fetch-feed": True}
    assert deps["use-bracket"] == {"fetch-feed": True}
    assert deps["ignored"] == {}                          # `with` scanned only for engram/executeStory (dag.go:3061-3066)


# ---- config 1 of BASELINE.json: A -> B -> C lifecycle
def _chain(**kw):
    return O.Story(steps=[O.Step("A", ref=True), O.Step("B", ref=True, needs=["A"]), O.Step("C", ref=True, needs=["B"])], **kw)


def _states(a, b, c, queued_b=False):
    out = {}
    for n, p in (("A", a), ("B", b), ("C", c)):
        if p:
            out[n] = O.StepState(p, message=(O.QUEUED_PREFIXES[0] + " (1 running, limit 1)") if (n == "B" and queued_b) else "")
    return out


@pytest.mark.parametrize("phases,ff,ready,skip", [
    (("", "", ""), True, ["A"], []),
    (("Running", "", ""), True, [], []),
    (("Succeeded", "", ""), True, ["B"], []),
    (("Succeeded", "Succeeded", ""), True, ["C"], []),
    (("Failed", "", ""), False, [], ["B"]),
    (("Failed", "Skipped", ""), False, ["C"], []),          # skipping is not transitive (dag.go:3377)
    (("Succeeded", "Blocked", ""), True, ["B"], []),         # Blocked/Scheduling/"" are in no set (dag.go:3377-3388)
    (("Succeeded", "Scheduling", ""), True, ["B"], []),
])
def test_config1_chain_snapshots(phases, ff, ready, skip):
    story = _chain(continue_on_step_failure=(None if ff else True))
    srun = O.StoryRun(step_states=_states(*phases))
    it = O.run_dag_iteration(srun, story)
    assert it.ready.ready == ready and it.ready.skipped == skip
    got = _packed_pass(story, O.StoryRun(step_states=_states(*phases)))
    assert got["ready"] == ready and got["skip"] == skip


def test_config1_all_succeeded_finalizes():
    story = _chain()
    it = O.run_dag_iteration(O.StoryRun(step_states=_states("Succeeded", "Succeeded", "Succeeded")), story)
    assert it.group == "finalize" and it.main_done
    got = _packed_pass(story, O.StoryRun(step_states=_states("Succeeded", "Succeeded", "Succeeded")))
    assert got["summary"] & A.SUM_GROUP_MASK == A.GROUP_DONE and got["summary"] & A.SUM_MAIN_DONE


def test_config1_fail_fast_marks_skipped():
    story = _chain()
    srun = O.StoryRun(step_states=_states("Failed", "", ""))
    it = O.run_dag_iteration(srun, story)
    assert srun.step_states["B"].phase == "Skipped" and srun.step_states["C"].phase == "Skipped"
    assert srun.step_states["B"].message == "Skipped due to fail-fast policy"
    assert it.group == "finalize" and it.main_failed
    got = _packed_pass(story, O.StoryRun(step_states=_states("Failed", "", "")))
    assert got["phase"] == {"A": "Failed", "B": "Skipped", "C": "Skipped"}
    assert got["summary"] & A.SUM_MAIN_FAILED and got["summary"] & A.SUM_GROUP_MASK == A.GROUP_DONE


def test_config1_allow_failure_counts_as_completed():
    story = _chain()
    story.steps[0].allow_failure = True
    it = O.run_dag_iteration(O.StoryRun(step_states=_states("Failed", "", "")), story)
    assert it.ready.ready == ["B"]
    assert _packed_pass(story, O.StoryRun(step_states=_states("Failed", "", "")))["ready"] == ["B"]


def test_config1_queued_pending_is_a_candidate():
    story = _chain()
    it = O.run_dag_iteration(O.StoryRun(step_states=_states("Succeeded", "Pending", "", queued_b=True)), story)
    assert it.ready.ready == ["B"]       # queued bit removes B from running (dag.go:2020-2033)
    assert _packed_pass(story, O.StoryRun(step_states=_states("Succeeded", "Pending", "", queued_b=True)))["ready"] == ["B"]
    it = O.run_dag_iteration(O.StoryRun(step_states=_states("Succeeded", "Pending", "")), story)
    assert it.ready.ready == []          # plain Pending is running
    assert _packed_pass(story, O.StoryRun(step_states=_states("Succeeded", "Pending", "")))["ready"] == []


def test_config1_realtime_running_dep_satisfies():
    story = _chain(realtime=True)
    it = O.run_dag_iteration(O.StoryRun(step_states=_states("Running", "", "")), story)
    assert it.ready.ready == ["B"]
    assert _packed_pass(story, O.StoryRun(step_states=_states("Running", "", "")))["ready"] == ["B"]


def test_compensation_group_allows_failed_dependencies():
    story = O.Story(steps=[O.Step("work", ref=True)],
                    compensations=[O.Step("undo-1", ref=True), O.Step("undo-2", ref=True, needs=["undo-1"])])
    st = {"work": O.StepState("Failed"), "undo-1": O.StepState("Failed")}
    it = O.run_dag_iteration(O.StoryRun(step_states=dict(st)), story)
    assert it.group == "compensation" and it.ready.ready == ["undo-2"]   # allowFailed (dag.go:2722)
    assert _packed_pass(story, O.StoryRun(step_states=dict(st)))["ready"] == ["undo-2"]


def test_mixed_failed_and_unmet_dependency_bounds():
    """X needs [F, U]; Go's map order decides — oracle exposes both bounds, contract = skip_max."""
    story = O.Story(steps=[O.Step("F", ref=True), O.Step("U", ref=True), O.Step("X", ref=True, needs=["F", "U"])],
                    continue_on_step_failure=True)
    srun = O.StoryRun(step_states={"F": O.StepState("Failed"), "U": O.StepState("Running")})
    it = O.run_dag_iteration(srun, story)
    assert it.ready.skipped == ["X"] and it.ready_min.skipped == []
    assert it.ready.ready == it.ready_min.ready == []
    got = _packed_pass(story, O.StoryRun(step_states={"F": O.StepState("Failed"), "U": O.StepState("Running")}))
    assert got["skip"] == ["X"]


def test_template_safety_violation_fails_step():
    story = O.Story(steps=[O.Step("danger", ref=True, if_='{{ env "HOME" }}')])
    srun = O.StoryRun()
    it = O.run_dag_iteration(srun, story, evaluator=O.literal_evaluator)
    assert it.ready.ready == [] and it.ready.failed_now == ["danger"]
    assert srun.step_states["danger"].phase == "Failed"
    got = _packed_pass(story, O.StoryRun(), evaluator=O.literal_evaluator)
    assert got["fail"] == ["danger"] and got["phase"]["danger"] == "Failed" and got["needs_cond"] == ["danger"]


def test_fixpoint_runs_primitives_to_quiescence():
    """runDagIterations (dag.go:393-540): condition steps complete instantly and unblock dependents."""
    story = O.Story(steps=[O.Step("c1", type="condition"), O.Step("c2", type="condition", needs=["c1"]),
                           O.Step("work", ref=True, needs=["c2"]), O.Step("after", ref=True, needs=["work"])])
    srun = O.StoryRun()
    iters, launched, skipped, _, last = O.run_dag_iterations(srun, story)
    assert launched == ["c1", "c2", "work"] and skipped == []
    assert srun.step_states["work"].phase == "Running" and "after" not in srun.step_states
    got = _packed_pass(story, O.StoryRun(), flags=A.EVAL_FIXPOINT)
    assert got["ready"] == ["c1", "c2", "work"]
    assert got["phase"] == {"c1": "Succeeded", "c2": "Succeeded", "work": "Running", "after": ""}
    assert (got["summary"] >> A.SUM_ITER_SHIFT) == iters


def test_parallel_expansion_count():
    step = O.Step("fan", type="parallel", with_={"steps": [{"name": "b%d" % i} for i in range(5)]})
    story = O.Story(steps=[step])
    got = _packed_pass(story, O.StoryRun())
    assert got["ready"] == ["fan"] and got["n_expansion"] == 5 and got["counts"]["expansion"] == 5
