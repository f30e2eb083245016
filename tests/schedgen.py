"""Random clusters for the limiter tests: the same snapshot as reference-shaped objects (StoryRun / StepRun lists, the
scheduling config) and as the packed bf_schedule inputs."""
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from oracle import limiters as LM
from oracle.pyoracle import StepState

RUN_PHASES = ["", "Pending", "Running", "Succeeded", "Failed", "Paused", "Finished"]
RUN_PHASE_CODE = {"": 0, "Pending": 1, "Running": 2, "Succeeded": 3, "Failed": 4, "Finished": 5, "Paused": 8}


@dataclass
class Cluster:
    cfg: LM.SchedulingConfig
    now: float
    runs: List[LM.ClusterStoryRun]
    step_runs: List[LM.ClusterStepRun]
    story_of_run: List[str]
    queue_of_run: List[str]
    n_steps: np.ndarray
    sched: np.ndarray
    ready_masks: np.ndarray
    run_running: np.ndarray
    run_demand: np.ndarray
    story_limit: np.ndarray
    story_base: np.ndarray
    queue_limit: np.ndarray
    queue_aging: np.ndarray
    queue_base: np.ndarray
    global_base: int
    step_phase: List[np.ndarray]   # per run: phase code per step (2 Running, 14 queued, else anything)

    def mask_names(self, mask) -> List[str]:
        return ["s%d" % (32 * w + b) for w in range(len(mask)) for b in range(32) if (int(mask[w]) >> b) & 1]

    def ready_names(self, r) -> List[str]:
        return self.mask_names(self.ready_masks[r])


def random_cluster(rng: np.random.Generator, n_runs: int = 0, s_max: int = 70) -> Cluster:
    now = 50_000.0
    queue_names = ["default"] + ["q%d" % i for i in range(int(rng.integers(0, 4)))]
    cfg = LM.SchedulingConfig(global_concurrency=int(rng.choice([0, 0, 3, 8, 25])), queues={})
    for qn in queue_names:
        if rng.random() < 0.8:   # a queue without an entry reads as the zero QueueConfig (scheduling.go:101-112)
            cfg.queues[qn] = LM.QueueConfig(int(rng.choice([0, 0, 2, 5, 12])), 0, int(rng.choice([0, 30, 60, 600])))
    n_stories = int(rng.integers(1, 7))
    stories = []
    for i in range(n_stories):
        qn = str(rng.choice(queue_names))
        stories.append(dict(name="story%d" % i, ns=str(rng.choice(["ns", "ns2"])), limit=int(rng.choice([0, 0, 1, 2, 4, 9])),
                            queue=("" if qn == "default" and rng.random() < 0.5 else qn), priority=int(rng.integers(-2, 6))))
    n = n_runs or int(rng.integers(3, 40))
    runs, step_runs, story_of_run, queue_of_run, phases = [], [], [], [], []
    sched = np.zeros(n, dtype=LM.SCHED_RUN_DTYPE)
    n_steps = rng.integers(1, s_max + 1, size=n)
    W = (int(n_steps.max()) + 31) // 32
    ready = np.zeros((n, W), dtype=np.uint32)
    run_running = np.zeros(n, dtype=np.uint32)
    run_demand = np.zeros(n, dtype=np.uint32)
    for r in range(n):
        st = stories[int(rng.integers(0, n_stories))]
        qname = st["queue"] or "default"
        qlabel = LM.queue_label_value(st["queue"])
        S = int(n_steps[r])
        states: Dict[str, StepState] = {}
        ph = np.zeros(S, dtype=np.uint8)
        earliest = None
        for i in range(S):
            u = rng.random()
            if u < 0.12:
                states["s%d" % i] = StepState("Running")
                ph[i] = 2
                step_runs.append(LM.ClusterStepRun(st["ns"], st["name"], qlabel, "Running"))
            elif u < 0.22:
                t = None if rng.random() < 0.2 else now - float(rng.integers(-5, 400)) - float(rng.random())
                states["s%d" % i] = StepState("Pending", str(rng.choice(LM.QUEUED_PREFIXES)) + " (1 running, limit 1)", started_at=t)
                ph[i] = 14
                if t is not None and (earliest is None or t < earliest):
                    earliest = t
            elif u < 0.30:
                states["s%d" % i] = StepState("Pending", "waiting")
                ph[i] = 1
            elif u < 0.55:
                states["s%d" % i] = StepState("Succeeded")
                ph[i] = 3
                if rng.random() < 0.3:
                    step_runs.append(LM.ClusterStepRun(st["ns"], st["name"], qlabel, "Succeeded"))
            elif rng.random() < 0.5:
                ready[r, i >> 5] |= np.uint32(1 << (i & 31))
        phase = str(rng.choice(RUN_PHASES))
        runs.append(LM.ClusterStoryRun("run%d" % r, st["ns"], qlabel, str(st["priority"]), phase, states))
        story_of_run.append(st["name"])
        queue_of_run.append(st["queue"])
        phases.append(ph)
        sched["story_key"][r] = stories.index(st)
        sched["queue_key"][r] = queue_names.index(qname)
        sched["priority"][r] = st["priority"]
        sched["queued_elapsed_s"][r] = LM.NONE_U32 if earliest is None else max(0, int(np.floor(now - earliest)))
        sched["run_phase"][r] = RUN_PHASE_CODE[phase]
        run_running[r] = int((ph == 2).sum())
        run_demand[r] = int(((ph == 2) | (ph == 14)).any())
    # StepRuns the batch does not account for: other runs of the batch's stories, and foreign stories
    story_base = np.zeros(n_stories, dtype=np.uint32)
    queue_base = np.zeros(len(queue_names), dtype=np.uint32)
    global_base = 0
    for _ in range(int(rng.integers(0, 12))):
        if rng.random() < 0.6:
            k = int(rng.integers(0, n_stories))
            st = stories[k]
            step_runs.append(LM.ClusterStepRun(st["ns"], st["name"], LM.queue_label_value(st["queue"]), "Running"))
            story_base[k] += 1
            queue_base[queue_names.index(st["queue"] or "default")] += 1
        else:
            q = int(rng.integers(0, len(queue_names)))
            step_runs.append(LM.ClusterStepRun("elsewhere", "foreign", queue_names[q], "Running"))
            queue_base[q] += 1
        global_base += 1
    order = rng.permutation(len(step_runs))
    step_runs = [step_runs[i] for i in order]
    story_limit = np.array([s["limit"] for s in stories], dtype=np.int32)
    queue_limit = np.array([LM.queue_config_for(cfg, q).concurrency for q in queue_names], dtype=np.int32)
    queue_aging = np.array([LM.queue_config_for(cfg, q).priority_aging_seconds for q in queue_names], dtype=np.int32)
    return Cluster(cfg, now, runs, step_runs, story_of_run, queue_of_run, n_steps, sched, ready, run_running, run_demand,
                   story_limit, story_base, queue_limit, queue_aging, queue_base, global_base, phases)
