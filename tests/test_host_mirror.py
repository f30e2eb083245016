"""C++ host mirror (include/bobrafrontier_host.h) vs the oracle's own object handling.

CPU part: the template-reference scanner against the reference's regex (dag.go:3028-3030, run by Python's re),
CSR / flags / state records against the test-side numpy packer, result decoding.  GPU part: Story objects ->
bfh_* -> bf_eval -> step names, against pyoracle."""
import ctypes as C
import random

import numpy as np
import pytest

from bobrapet_b200 import _abi as A
from bobrapet_b200 import host as H
from bobrapet_b200.records import unpack_result
from oracle import packed as PK
from oracle import pyoracle as O
from tests import packing as P
from tests.test_oracle_differential import _evaluator, _rand_run, _rand_story

FRAGS = ["steps.", "steps", ".", "a-b", "x_1", "[", "]", " ", "'", '"', "(index", " .steps ", ")", "steps.a.", "steps['b']",
         '(index .steps "c")', "{{", "}}", "steps .q.", "steps\t[ \"w\" ]", "steps.a", "(index  .steps\t'z-1')", "(index .steps \"n\" )"]


def _py_refs(expr):
    return [m.group(3) or m.group(2) or m.group(1) for m in O.STEP_NAME_REGEX.finditer(expr)]


def test_scanner_matches_reference_regex():
    rng = random.Random(7)
    for _ in range(3000):
        expr = "".join(rng.choice(FRAGS) for _ in range(rng.randint(1, 12)))
        assert H.scan_step_refs(expr) == _py_refs(expr), expr
    for expr in ["", "steps.", "steps..", "steps.a.steps.b.", "xsteps.a.", "steps.a.b.c.", "steps[ 'a' ]steps[\"b\"]"]:
        assert H.scan_step_refs(expr) == _py_refs(expr), expr


def _host_story(story: O.Story) -> H.HostStory:
    hs = H.HostStory()
    for grp, steps in ((A.GROUP_MAIN, story.steps), (A.GROUP_COMPENSATION, story.compensations), (A.GROUP_FINALLY, story.finally_)):
        for st in steps:
            t = A.STEP_ENGRAM if st.ref else A.STEP_TYPE_CODE[st.type]
            skip = st.type in ("gate", "wait") and isinstance(st.with_, dict) and str(st.with_.get("onTimeout", "")).lower() == "skip"
            branches = [(b.name, bool(b.allow_failure)) for b in O.parse_parallel_branches(st)] if (st.type == "parallel" and st.with_) else []
            hs.add_step(st.name, grp, t, bool(st.allow_failure), skip, st.if_, (st.with_raw() if st.with_ is not None else None),
                        st.needs, branches)
    hs.set_policy(story.continue_on_step_failure, story.realtime)
    return hs


@pytest.mark.parametrize("seed", range(150))
def test_story_packing_matches_python_packer(seed):
    rng = random.Random(seed)
    story, codes = _rand_story(rng)
    hs = _host_story(story)
    rc = hs.finalize()
    err = O.validate_runtime_dependency_graph(O.all_story_steps(story))
    try:
        ps = P.pack_story(story)
    except ValueError:
        assert rc == A.BF_ETOPO and "unknown step dependencies" in hs.error()
        return
    assert rc == 0, hs.error()
    rp, ci, fl, nP = hs.csr()
    assert np.array_equal(rp, ps.row_ptr) and np.array_equal(ci, ps.col_idx) and np.array_equal(fl, ps.flags)
    assert nP == len(ps.par_steps)
    assert [hs.name(i) for i in range(ps.S)] == ps.names and all(hs.index(n) == i for i, n in enumerate(ps.names))
    assert hs.run_flags() == P.run_flags_of(story, O.StoryRun())
    if err is not None:
        assert "cycle" in err  # cycles are rejected at bf_topology_put (needs a device); unknown deps here

    # ---- a run: the in-place state record equals the numpy packer's
    srun, step_runs = _rand_run(rng, story)
    ev, now = _evaluator(codes), 100.0
    L, want = P.pack_runs([story], [ps], [srun], [0], [5], [step_runs], ev, {"inputs": {}, "steps": {}}, now)
    hb = H.HostBatch(None, ps.S, ps.child_nibbles(), L.fields, 4)
    assert hb.L.as_dict() == L.as_dict()
    r = hb.add_run(hs, 5)
    all_steps = O.all_story_steps(story)
    for i, st in enumerate(all_steps):
        ss = srun.step_states.get(st.name)
        if ss is not None:
            hb.set_phase(r, i, ss.phase, ss.message)
        hb.set_cond(r, i, P.cond_code(st, story, ev, {"inputs": {}, "steps": {}}))
        if st.type == "gate" and not st.ref:
            gs = srun.gates.get(st.name)
            timeout, _, _ = O.parse_gate_config(st)
            started = ss.started_at if (ss and ss.started_at is not None) else now
            hb.set_gate(r, i, gs.state if gs else "", timeout is not None and not (now < started + timeout))
        elif st.type in ("sleep", "wait") and not st.ref:
            hb.set_decision(r, i, P.decision_code(st, srun, now, None, ev, {"inputs": {}, "steps": {}}))
    hb.set_run_flags(r, srun.topology_terminated, -1)
    srs = {sr.name: sr for sr in step_runs}
    for q, stp in enumerate(ps.par_steps):
        kids = srun.primitive_children.get(ps.names[stp]) or []
        if not kids:
            continue
        hb.register_children(r, q, True)
        by_id = {srs[k].step_id: srs[k] for k in kids if k in srs}
        for b, bn in enumerate(ps.par_branches[q]):
            if bn in by_id:
                hb.set_child_phase(r, q, b, by_id[bn].phase)
    assert np.array_equal(hb.state(), want), (np.nonzero(hb.state()[0] != want[0])[0])

    # ---- results: decode what the oracle produced for this record
    res, _ = PK.evaluate(PK.PackedTopologies(P.topology_set([ps]), [5]), L, want)
    C.memmove(hb._l.bfh_batch_result(hb._p), res.ctypes.data, res.nbytes)
    out = unpack_result(L, res, ps.S)
    assert hb.ready(r) == np.nonzero(out["ready"][0])[0].tolist()
    assert hb.skipped(r) == np.nonzero(out["skip"][0])[0].tolist()
    assert hb.failed(r) == np.nonzero(out["fail"][0])[0].tolist()
    assert hb.needs_cond(r) == np.nonzero(out["needs_cond"][0])[0].tolist()
    assert [hb.phase_out(r, i) for i in range(ps.S)] == out["phase_out"][0].tolist()
    assert hb.summary(r) == int(out["summary"][0])
    it = O.run_dag_iteration(srun, story, step_runs, ev, {"inputs": {}, "steps": {}}, now)
    for name in it.ready.skipped:
        reason = hb.skip_reason(r, ps.index[name])
        want_reason = it.ready.skip_reasons[name]
        if "failed dependency" in want_reason:  # Go names a random failed dep (map order); the mirror names the first in CSR order
            assert reason.startswith("Skipped due to failed dependency: ")
            named = reason.split(": ")[1]
            ds = srun.step_states.get(named)
            assert named in O.build_dependency_graphs(
                {"main": story.steps, "compensation": story.compensations, "finally": story.finally_}[it.group])[0][name]
            assert ds is not None and O.is_terminal(ds.phase) and ds.phase not in ("Succeeded", "Skipped")
        else:
            assert reason == want_reason
    hb.close()
    hs.close()


def test_unknown_dependency_message_matches_reference_format():
    hs = H.HostStory()
    hs.add_step("a", needs=["ghost"])
    assert hs.finalize() == A.BF_ETOPO
    assert hs.error() == "unknown step dependencies: a->ghost"   # dag_test.go:206, dag.go:3097


def test_symbols_exported():
    lib = A.load()
    for n in H.HOST_SYMBOLS:
        assert hasattr(lib, n), n


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(60))
def test_end_to_end_through_host_mirror(seed):
    from bobrapet_b200 import Frontier
    rng = random.Random(9000 + seed)
    story, codes = _rand_story(rng)
    if O.validate_runtime_dependency_graph(O.all_story_steps(story)) is not None:
        pytest.skip("invalid graph")
    hs = _host_story(story)
    if hs.finalize() != 0:
        pytest.skip("dangling alias")
    fr = Frontier(0)
    try:
        slot = hs.upload(fr)
        ps = P.pack_story(story)
        hb = H.HostBatch(fr, ps.S, ps.child_nibbles(), A.F_COND | A.F_DECISION | A.F_ALL_OUT | (A.F_CHILD if ps.child_nibbles() else 0), 8)
        srun, step_runs = _rand_run(rng, story)
        ev, now = _evaluator(codes), 100.0
        L, want_state = P.pack_runs([story], [ps], [srun], [0], [slot], [step_runs], ev, {"inputs": {}, "steps": {}}, now)
        r = hb.add_run(hs, slot)
        C.memmove(hb._l.bfh_batch_state(hb._p), want_state.ctypes.data, want_state.nbytes)  # state packing is covered on CPU
        hb.eval(A.EVAL_VALIDATE)
        it = O.run_dag_iteration(srun, story, step_runs, ev, {"inputs": {}, "steps": {}}, now)
        order = {n: i for i, n in enumerate(ps.names)}
        assert [hs.name(i) for i in hb.ready(r)] == sorted(it.ready.ready, key=order.get)
        assert [hs.name(i) for i in hb.skipped(r)] == sorted(it.ready.skipped, key=order.get)
        assert [hb.phase_out(r, i) for i in range(ps.S)] == [P.phase_code(srun.step_states.get(n)) for n in ps.names]
        hb.close()
    finally:
        hs.close()
        fr.close()
