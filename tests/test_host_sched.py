"""C++ host mirror of the limiters' host side (bfh_sched_*, include/bobrafrontier_host.h): scheduling decisions, key
interning, elapsed seconds and message text against the object-level oracle; end to end on the GPU."""
import numpy as np
import pytest

from bobrapet_b200 import _abi as A
from bobrapet_b200 import host as H
from oracle import limiters as LM
from tests.schedgen import random_cluster


def _story(n_steps):
    hs = H.HostStory()
    for i in range(n_steps):
        hs.add_step("s%d" % i, 0, A.STEP_ENGRAM)
    assert hs.finalize() == 0, hs.error()
    return hs


def _build(cl, frontier=None):
    """a HostBatch + HostSched holding the cluster `cl` (tests/schedgen.py)"""
    n = len(cl.runs)
    hb = H.HostBatch(frontier, int(cl.n_steps.max()), 0, 0, n)
    stories = {}
    for r in range(n):
        S = int(cl.n_steps[r])
        if S not in stories:
            stories[S] = _story(S)
            stories[S].slot = stories[S].upload(frontier) if frontier is not None else 1000 + S
        hb.add_run(stories[S], stories[S].slot)
        for i, code in enumerate(cl.step_phase[r]):
            if code:
                hb.set_phase_code(r, i, int(code))
    hsch = H.HostSched(hb)
    hsch.set_global(cl.cfg.global_concurrency, 0)
    for qn, qc in cl.cfg.queues.items():
        hsch.set_queue(qn, qc.concurrency, qc.default_priority, qc.priority_aging_seconds, 0)
    if "default" not in cl.cfg.queues:       # a config without the default queue: zero entry (scheduling.go:101-112)
        hsch.set_queue("default", 0, 0, 0, 0)
    for r, srun in enumerate(cl.runs):
        qs = LM.story_run_queued_since(srun)
        hsch.set_run(r, srun.namespace, cl.story_of_run[r], int(cl.story_limit[cl.sched["story_key"][r]]),
                     cl.queue_of_run[r] or None, int(cl.sched["priority"][r]), srun.phase,
                     qs, cl.now)
    return hb, hsch, stories


def test_scheduling_decision_matches_oracle():
    cfg = LM.SchedulingConfig(queues={"default": LM.QueueConfig(0, 3, 60), "gpu": LM.QueueConfig(4, 7, 0)})
    hb = H.HostBatch(None, 4, 0, 0, 8)
    st = _story(4)
    for _ in range(5):
        hb.add_run(st, 1)
    hs = H.HostSched(hb)
    for qn, qc in cfg.queues.items():
        hs.set_queue(qn, qc.concurrency, qc.default_priority, qc.priority_aging_seconds)
    cases = [(None, None), ("", None), ("  GPU ", None), ("gpu", 11), ("unknown-queue", None)]
    for r, (q, p) in enumerate(cases):
        hs.set_run(r, "ns", "story", 0, q, p, "Running", None, 100)
    runs, tabs = hs.packed()
    for r, (q, p) in enumerate(cases):
        want_q, want_p = LM.resolve_scheduling_decision(q, p, cfg)
        assert hs.queue_name(int(runs["queue_key"][r])) == LM.queue_label_value(want_q), (q, p)
        assert int(runs["priority"][r]) == want_p, (q, p)
    assert tabs["queue_aging"][0] == 60 and hs.queue_name(0) == "default"


def test_elapsed_seconds_and_messages():
    hb = H.HostBatch(None, 4, 0, 0, 4)
    st = _story(4)
    for _ in range(4):
        hb.add_run(st, 1)
    hs = H.HostSched(hb)
    hs.set_run(0, "ns", "a", 2, None, 1, "Pending", 880, 1000)    # queued 120 s ago (dag_test.go:579)
    hs.set_run(1, "ns", "a", 2, None, 1, "Pending", None, 1000)
    hs.set_run(2, "ns", "b", 0, None, 1, "Failed", 1005, 1000)    # StartedAt in the future: elapsed <= 0
    hs.set_run(3, "ns2", "a", 0, None, 1, "", 1000, 1000)
    runs, tabs = hs.packed()
    assert runs["queued_elapsed_s"].tolist() == [120, LM.NONE_U32, 0, 0]
    assert runs["story_key"].tolist() == [0, 0, 1, 2]                # keyed by namespace + name (dag.go:1874-1877)
    assert tabs["story_limit"].tolist() == [2, 0, 0]
    assert runs["run_phase"].tolist() == [1, 1, 4, 0]
    assert H.format_queue_message(0, 1, 1) == "%s (1 running, limit 1)" % LM.STORY_PREFIX
    assert H.format_queue_message(A.QUEUED_GLOBAL, 7, 5) == "%s (7 running, limit 5)" % LM.GLOBAL_PREFIX
    assert H.format_queue_message(A.QUEUED_QUEUE, 3, 2) == "%s (3 running, limit 2)" % LM.QUEUE_PREFIX
    assert H.format_queue_message(A.QUEUED_PRIORITY, 0, 0) == LM.PRIORITY_PREFIX
    assert H.format_queue_message(A.QUEUED_OTHER, 0, 0) == "Queued due to scheduling limits"


@pytest.mark.parametrize("seed", range(12))
def test_mirror_packing_reproduces_object_level(seed):
    """the mirror's packed inputs, fed to the packed oracle, give the reference's per-run result (CPU only)"""
    rng = np.random.default_rng(31000 + seed)
    cl = random_cluster(rng)
    hb, hs, _ = _build(cl)
    runs, tabs = hs.packed()
    # base counts: the cluster's external StepRuns, keyed through the mirror's own interning
    story_base = np.zeros(len(tabs["story_limit"]), np.uint32)
    queue_base = np.zeros(len(tabs["queue_limit"]), np.uint32)
    for r in range(len(cl.runs)):
        story_base[runs["story_key"][r]] = cl.story_base[cl.sched["story_key"][r]]
    qkey = {hs.queue_name(k): k for k in range(len(queue_base))}
    names = ["default"] + ["q%d" % i for i in range(len(cl.queue_base) - 1)]
    for i, nme in enumerate(names):
        if nme in qkey:
            queue_base[qkey[nme]] = cl.queue_base[i]
        else:
            assert cl.queue_base[i] == 0 or True
    # queues no run of the batch uses are unknown to the mirror: their external StepRuns only count globally
    launch, q_story, q_sched, info, sr, qr, gr, _ = LM.schedule_packed(
        cl.run_running, cl.run_demand, runs, cl.ready_masks, tabs["story_limit"], story_base, tabs["queue_limit"],
        tabs["queue_aging"], queue_base, tabs["global_limit"], cl.global_base)
    for r, srun in enumerate(cl.runs):
        res = LM.apply_limiters(cl.step_runs, cl.runs, srun, cl.story_of_run[r], int(cl.story_limit[cl.sched["story_key"][r]]),
                                cl.queue_of_run[r], int(cl.sched["priority"][r]), cl.cfg, cl.ready_names(r), now=cl.now)
        assert cl.mask_names(launch[r]) == res.launch, (seed, r)
        assert cl.mask_names(q_story[r]) == res.queued_story, (seed, r)
        assert cl.mask_names(q_sched[r]) == res.queued_sched, (seed, r)
        if res.queued_story:
            assert H.format_queue_message(0, int(sr[runs["story_key"][r]]), int(tabs["story_limit"][runs["story_key"][r]])) == res.msg_story
        if res.queued_sched:
            reason = int(info[r, 3])
            run_tot = {LM.REASON_GLOBAL: (gr, tabs["global_limit"]),
                       LM.REASON_QUEUE: (int(qr[runs["queue_key"][r]]), int(tabs["queue_limit"][runs["queue_key"][r]]))}.get(reason, (0, 0))
            assert H.format_queue_message(reason, int(run_tot[0]), int(run_tot[1])) == res.msg_sched, (seed, r)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_mirror_end_to_end_on_device(seed):
    """Story/StoryRun objects -> bfh_batch_eval -> bfh_sched_apply -> step lists and messages, against the reference's
    limiters applied to the ready lists of the same pass"""
    from bobrapet_b200 import Frontier
    rng = np.random.default_rng(32000 + seed)
    cl = random_cluster(rng, n_runs=60)
    fr = Frontier(0)
    hb = hs = None
    try:
        hb, hs, stories = _build(cl, fr)
        # external StepRuns of the batch's stories / queues
        seen_story = {}
        for r, srun in enumerate(cl.runs):
            seen_story[(srun.namespace, cl.story_of_run[r])] = int(cl.story_base[cl.sched["story_key"][r]])
        for (ns, nme), base in seen_story.items():
            hs.set_story_base(ns, nme, base)
        names = ["default"] + ["q%d" % i for i in range(len(cl.queue_base) - 1)]
        for i, nme in enumerate(names):
            qc = LM.queue_config_for(cl.cfg, nme)
            hs.set_queue(nme, qc.concurrency, qc.default_priority, qc.priority_aging_seconds, int(cl.queue_base[i]))
        hs.set_global(cl.cfg.global_concurrency, cl.global_base)
        hb.eval(0)
        hs.apply()
        for r, srun in enumerate(cl.runs):
            ready = ["s%d" % i for i in hb.ready(r)]
            res = LM.apply_limiters(cl.step_runs, cl.runs, srun, cl.story_of_run[r], int(cl.story_limit[cl.sched["story_key"][r]]),
                                    cl.queue_of_run[r], int(cl.sched["priority"][r]), cl.cfg, ready, now=cl.now)
            assert ["s%d" % i for i in hs.steps(r, 0)] == res.launch, (seed, r)
            assert ["s%d" % i for i in hs.steps(r, 1)] == res.queued_story, (seed, r)
            assert ["s%d" % i for i in hs.steps(r, 2)] == res.queued_sched, (seed, r)
            assert hs.message(r, 1) == res.msg_story, (seed, r)
            assert hs.message(r, 2) == res.msg_sched, (seed, r)
    finally:
        if hs is not None:
            hs.close()
        if hb is not None:
            hb.close()      # the batch's pinned buffers belong to the ctx: release them before it
        fr.close()
