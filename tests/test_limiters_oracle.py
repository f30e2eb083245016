"""Limiter oracle (row a9 / f4): pinned by the reference's own known-answer tests, then the packed batch
contract (oracle.limiters.schedule_packed == bf_schedule) is checked against the object level on random clusters."""
import numpy as np
import pytest

from oracle import limiters as LM
from oracle.pyoracle import StepState
from tests.schedgen import random_cluster


def test_kat_priority_ordering_blocks_lower_priority():
    """dag_test.go:528-575: high (prio 10, Running) blocks low (prio 1, Pending) in the same queue."""
    high = LM.ClusterStoryRun("high", "default", "default", "10", "Running")
    low = LM.ClusterStoryRun("low", "default", "default", "1", "Pending")
    blocked, reason = LM.enforce_priority_ordering([high, low], low, "default", "default", 1, LM.SchedulingConfig(), now=1000.0)
    assert blocked and reason != ""
    # the high-priority run itself is not blocked by the low one
    blocked, _ = LM.enforce_priority_ordering([high, low], high, "default", "default", 10, LM.SchedulingConfig(), now=1000.0)
    assert not blocked


def test_kat_priority_ordering_allows_aged_run():
    """dag_test.go:579-633: low (prio 1) queued for 2 minutes under the default 60 s aging reaches 3 > 2."""
    now = 10_000.0
    high = LM.ClusterStoryRun("high", "default", "default", "2", "Running")
    low = LM.ClusterStoryRun("low", "default", "default", "1", "Pending",
                             {"queued": StepState("Pending", LM.PRIORITY_PREFIX, started_at=now - 120.0)})
    blocked, _ = LM.enforce_priority_ordering([high, low], low, "default", "default", 1, LM.SchedulingConfig(), now=now)
    assert not blocked
    assert LM.effective_priority(1, now - 120.0, 60, now) == 3
    assert LM.effective_priority(1, now - 59.0, 60, now) == 1        # int32(elapsed.Seconds()) / aging == 0
    assert LM.effective_priority(1, None, 60, now) == 1
    assert LM.effective_priority(1, now + 5.0, 60, now) == 1         # elapsed <= 0
    assert LM.effective_priority(1, now - 500.0, 0, now) == 1        # aging disabled


def test_kat_story_concurrency_queues_everything():
    """dag_test.go:744-826: limit 1, one Running StepRun of the story -> ready 0, both steps queued with a message."""
    sr = [LM.ClusterStepRun("default", "story", "default", "Running")]
    srun = LM.ClusterStoryRun("srun", "default", "default", None, "")
    res = LM.apply_limiters(sr, [srun], srun, "story", 1, "", 0, LM.SchedulingConfig(), ["step-a", "step-b"], now=0.0)
    assert res.launch == [] and res.queued_story == ["step-a", "step-b"] and res.queued_sched == []
    assert res.msg_story.startswith(LM.STORY_PREFIX) and "(1 running, limit 1)" in res.msg_story
    # limit 2: one slot, the list PREFIX is kept (dag.go:1796-1798)
    res = LM.apply_limiters(sr, [srun], srun, "story", 2, "", 0, LM.SchedulingConfig(), ["step-a", "step-b"], now=0.0)
    assert res.launch == ["step-a"] and res.queued_story == ["step-b"]
    # no limit
    res = LM.apply_limiters(sr, [srun], srun, "story", 0, "", 0, LM.SchedulingConfig(), ["step-a", "step-b"], now=0.0)
    assert res.launch == ["step-a", "step-b"] and not res.queued_story


def test_kat_global_count_is_running_step_runs_only():
    """dag_test.go:136-152: countRunningStepRunsGlobal lists StepRuns by the field selector status.phase=Running — the
    global limiter sees Running StepRuns of every story and queue, and nothing else."""
    sr = [LM.ClusterStepRun("a", "s1", "default", "Running"), LM.ClusterStepRun("b", "s2", "gpu", "Running"),
          LM.ClusterStepRun("a", "s1", "default", "Succeeded"), LM.ClusterStepRun("a", "s1", "default", "Pending"),
          LM.ClusterStepRun("a", "s1", "default", "")]
    cfg = LM.SchedulingConfig(global_concurrency=3, queues={"default": LM.QueueConfig()})
    srun = LM.ClusterStoryRun("r", "a", "default", "0", "Running")
    ready, queued, reason = LM.enforce_scheduling_limits(sr, [srun], srun, "", 0, cfg, ["x", "y"], now=0.0)
    assert ready == ["x"] and queued == ["y"] and "(2 running, limit 3)" in reason


def test_scheduling_limit_reasons():
    """dag.go:1845-1859: which limit names the reason."""
    cfg = LM.SchedulingConfig(global_concurrency=3, queues={"default": LM.QueueConfig(2, 0, 0)})
    sr = [LM.ClusterStepRun("ns", "s", "default", "Running")] * 1 + [LM.ClusterStepRun("ns", "t", "other", "Running")] * 2
    srun = LM.ClusterStoryRun("r", "ns", "default", "0", "Running")
    ready, queued, reason = LM.enforce_scheduling_limits(sr, [srun], srun, "default", 0, cfg, ["a", "b", "c"], now=0.0)
    # global: 3 - 3 = 0 slots, queue: 2 - 1 = 1 slot -> 0 slots, global names it (gslots <= qslots)
    assert ready == [] and queued == ["a", "b", "c"] and reason.startswith(LM.GLOBAL_PREFIX)
    cfg.global_concurrency = 10
    ready, queued, reason = LM.enforce_scheduling_limits(sr, [srun], srun, "default", 0, cfg, ["a", "b", "c"], now=0.0)
    assert ready == ["a"] and queued == ["b", "c"] and reason.startswith(LM.QUEUE_PREFIX) and "(1 running, limit 2)" in reason


def test_truncate_mask_keeps_list_prefix():
    m = np.array([0b1011_0000, 0, 0b1, 0xFFFF_FFFF], dtype=np.uint32)
    assert LM.truncate_mask(m, 0).tolist() == [0, 0, 0, 0]
    assert LM.truncate_mask(m, 2).tolist() == [0b0011_0000, 0, 0, 0]
    assert LM.truncate_mask(m, 4).tolist() == [0b1011_0000, 0, 1, 0]
    assert LM.truncate_mask(m, 6).tolist() == [0b1011_0000, 0, 1, 0b11]
    assert LM.truncate_mask(m, 99).tolist() == m.tolist()


@pytest.mark.parametrize("seed", range(40))
def test_packed_contract_equals_object_level(seed):
    """random clusters: the batch contract (counts by reduction, then independent truncation) gives, run by run,
    what the reference's per-reconcile limiters give on the same snapshot"""
    rng = np.random.default_rng(9000 + seed)
    cl = random_cluster(rng)
    launch, q_story, q_sched, info, story_running, queue_running, global_running, _ = LM.schedule_packed(
        cl.run_running, cl.run_demand, cl.sched, cl.ready_masks, cl.story_limit, cl.story_base, cl.queue_limit,
        cl.queue_aging, cl.queue_base, cl.cfg.global_concurrency, cl.global_base)
    for r, srun in enumerate(cl.runs):
        names = cl.ready_names(r)
        res = LM.apply_limiters(cl.step_runs, cl.runs, srun, cl.story_of_run[r], int(cl.story_limit[cl.sched["story_key"][r]]),
                                cl.queue_of_run[r], int(cl.sched["priority"][r]), cl.cfg, names, now=cl.now)
        assert cl.mask_names(launch[r]) == res.launch, (seed, r)
        assert cl.mask_names(q_story[r]) == res.queued_story, (seed, r)
        assert cl.mask_names(q_sched[r]) == res.queued_sched, (seed, r)
        reason = int(info[r, 3])
        if res.queued_sched:
            want = {LM.PRIORITY_PREFIX: LM.REASON_PRIORITY, LM.GLOBAL_PREFIX: LM.REASON_GLOBAL, LM.QUEUE_PREFIX: LM.REASON_QUEUE}
            assert reason == next(v for k, v in want.items() if res.msg_sched.startswith(k)), (seed, r, res.msg_sched)
            if reason == LM.REASON_GLOBAL:
                assert "(%d running, limit %d)" % (global_running, cl.cfg.global_concurrency) in res.msg_sched
            if reason == LM.REASON_QUEUE:
                q = int(cl.sched["queue_key"][r])
                assert "(%d running, limit %d)" % (queue_running[q], cl.queue_limit[q]) in res.msg_sched
        else:
            assert reason == LM.REASON_NONE
        if res.queued_story:
            k = int(cl.sched["story_key"][r])
            assert "(%d running, limit %d)" % (story_running[k], cl.story_limit[k]) in res.msg_story
