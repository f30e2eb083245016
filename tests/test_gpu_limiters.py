"""GPU parity of the limiters (rows a9 / f4): bf_eval -> bf_schedule through the C ABI vs oracle.limiters.schedule_packed,
bit for bit, on adversarial random batches (all step types, parallel children, every phase code) and random
stories / queues / limits / priorities."""
import numpy as np
import pytest

from bobrapet_b200 import _abi as A
from bobrapet_b200 import Frontier
from bobrapet_b200.records import _unpack_planes
from oracle import limiters as LM
from tests import randgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fr():
    f = Frontier(0)
    yield f
    f.close()


def _running_and_demand(ts, topo, L, state):
    """the checker's own derivation of the per-run reductions from the packed state records"""
    n = state.shape[0]
    W = L.words
    S_off = np.concatenate(([0], np.cumsum(ts.S.astype(np.int64))))
    P_off = np.concatenate(([0], np.cumsum(ts.P.astype(np.int64))))
    cfs, _ = randgen.child_layout(ts)
    run_running = np.zeros(n, dtype=np.uint32)
    run_demand = np.zeros(n, dtype=np.uint32)
    codes = _unpack_planes(state[:, L.off_phase:L.off_phase + W * 16], W, 4, W * 32)
    codes[codes == 15] = 0
    for r in range(n):
        t = int(topo[r])
        S = int(ts.S[t])
        ph = codes[r, :S]
        types = ts.step_flags[S_off[t]:S_off[t] + S] & A.SF_TYPE_MASK
        cnt = int(((ph == 2) & (types == 0)).sum())
        if L.off_child != A.OFF_NONE and int(ts.P[t]):
            reg = int(np.ascontiguousarray(state[r, 8:16]).view("<u8")[0])
            nb = (L.child_nibbles + 1) // 2
            raw = state[r, L.off_child:L.off_child + nb]
            nib = np.empty(nb * 2, dtype=np.uint8)
            nib[0::2] = raw & 0xF
            nib[1::2] = raw >> 4
            nib[nib == 15] = 0
            for q in range(int(ts.P[t])):
                if (reg >> q) & 1:
                    B = int(ts.parallel["branches"][P_off[t] + q])
                    c0 = int(cfs[t][q])
                    cnt += int((nib[c0:c0 + B] == 2).sum())
        run_running[r] = cnt
        run_demand[r] = int(((ph == 2) | (ph == 14)).any())
    return run_running, run_demand


@pytest.mark.parametrize("seed", range(8))
def test_schedule_matches_oracle(fr, seed):
    rng = np.random.default_rng(7000 + seed)
    smax = [40, 257, 1024, 96, 600, 33, 8, 130][seed]
    ts = randgen.random_topologies(rng, 40, 1, smax, parallel=(seed % 2 == 0))
    slots = fr.put_topologies(ts)
    n = [3000, 2999, 700, 3001, 1200, 3003, 2000, 2500][seed]
    L, state, topo = randgen.random_state(rng, ts, slots, n, A.F_COND | A.F_DECISION, phase_mix=("any" if seed % 2 else "progress"))
    result, _ = fr.eval(L, state)
    W = L.words
    ready = np.ascontiguousarray(result[:, L.off_ready:L.off_ready + 4 * W]).view("<u4").reshape(n, W)
    n_stories, n_queues = int(rng.integers(1, 30)), int(rng.integers(1, 6))
    sched = np.zeros(n, dtype=LM.SCHED_RUN_DTYPE)
    sched["story_key"] = rng.integers(0, n_stories, size=n)
    sched["queue_key"] = rng.integers(0, n_queues, size=n)
    sched["priority"] = rng.integers(-3, 8, size=n)
    el = rng.integers(0, 4000, size=n).astype(np.uint32)
    el[rng.random(n) < 0.5] = LM.NONE_U32
    sched["queued_elapsed_s"] = el
    sched["run_phase"] = rng.choice([0, 1, 2, 3, 4, 8, 11], size=n)
    # limits sized to the batch so that every branch (no limit / some slots / zero slots) occurs
    story_limit = rng.choice([0, 0, 1, 3, 40, 200, 1000], size=n_stories).astype(np.int32)
    queue_limit = rng.choice([0, 50, 2000, 20000], size=n_queues).astype(np.int32)
    queue_aging = rng.choice([0, 30, 60, 600], size=n_queues).astype(np.int32)
    global_limit = int(rng.choice([0, 100, 5000, 100000]))
    story_base = rng.integers(0, 5, size=n_stories).astype(np.uint32)
    queue_base = rng.integers(0, 50, size=n_queues).astype(np.uint32)
    global_base = int(queue_base.sum())
    got = fr.schedule(L, n, sched, story_limit, queue_limit, queue_aging, global_limit, story_base, queue_base, global_base)
    run_running, run_demand = _running_and_demand(ts, topo, L, state)
    launch, q_story, q_sched, info, sr, qr, gr, mp = LM.schedule_packed(
        run_running, run_demand, sched, ready, story_limit, story_base, queue_limit, queue_aging, queue_base, global_limit, global_base)
    assert np.array_equal(got["story_running"], sr.astype(np.uint32))
    assert np.array_equal(got["queue_running"], qr.astype(np.uint32))
    assert got["global_running"] == gr
    assert np.array_equal(got["queue_max_priority"].astype(np.int64), mp)
    rec = got["records"]
    hdr = np.ascontiguousarray(rec[:, 0:16]).view("<u4").reshape(n, 4)
    masks = np.ascontiguousarray(rec[:, 16:16 + 12 * W]).view("<u4").reshape(n, 3, W)
    assert np.array_equal(masks[:, 0], launch), "launch masks differ"
    assert np.array_equal(masks[:, 1], q_story), "story-queued masks differ"
    assert np.array_equal(masks[:, 2], q_sched), "sched-queued masks differ"
    assert np.array_equal(hdr, info)
    assert not rec[:, 16 + 12 * W:].any()
    # the three masks partition the ready set (every ready step is launched or queued, dag.go:1709-1728)
    assert np.array_equal(masks[:, 0] | masks[:, 1] | masks[:, 2], ready)


def test_schedule_requires_preceding_eval(fr):
    ts = randgen.random_topologies(np.random.default_rng(1), 4, 5, 20, parallel=False)
    slots = fr.put_topologies(ts)
    L, state, _ = randgen.random_state(np.random.default_rng(2), ts, slots, 10, 0)
    fr.eval(L, state)
    sched = np.zeros(11, dtype=LM.SCHED_RUN_DTYPE)
    with pytest.raises(A.FrontierError):
        fr.schedule(L, 11, sched, [0], [0], [0])   # n_runs differs from the evaluated batch
