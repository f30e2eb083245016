"""Redrive closure (row f3, storyrun_controller.go:535-558): object-level oracle on hand-made stories, the packed
statement (a fixed point over the CSR) against it on random stories, and the device kernel against both."""
import random

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import packing as P
from tests.test_oracle_differential import _rand_story


def closure_packed(ps: P.PackedStory, start: int) -> np.ndarray:
    """the packed contract: start + every step OF THE SAME GROUP with a dependency already selected, to the fixed point"""
    S = ps.S
    group = (ps.flags.astype(np.uint32) >> 6) & 3
    sel = np.zeros(S, dtype=bool)
    sel[start] = True
    changed = True
    while changed:
        changed = False
        for i in range(S):
            if sel[i] or group[i] != group[start]:
                continue
            if sel[ps.col_idx[ps.row_ptr[i]:ps.row_ptr[i + 1]]].any():
                sel[i] = True
                changed = True
    W = (S + 31) // 32
    return np.packbits(np.concatenate([sel, np.zeros(W * 32 - S, dtype=bool)]), bitorder="little").view("<u4")


def test_chain_and_diamond():
    # a -> b -> c ; d independent (config 1 of BASELINE.json plus a bystander)
    st = O.Story([O.Step("a", ref=True), O.Step("b", ["a"], ref=True), O.Step("c", ["b"], ref=True), O.Step("d", ref=True)])
    assert set(O.resolve_redrive_from_step_set(st, "a")) == {"a", "b", "c"}
    assert set(O.resolve_redrive_from_step_set(st, "b")) == {"b", "c"}
    assert set(O.resolve_redrive_from_step_set(st, "d")) == {"d"}
    with pytest.raises(KeyError):
        O.resolve_redrive_from_step_set(st, "nope")          # "step %q not found", :541
    # template reference counts as a dependency (buildDependencyGraphs, dag.go:3056-3070)
    st = O.Story([O.Step("fetch-data", ref=True), O.Step("use", ref=True, if_="{{ steps.fetch_data.output.ok }}")])
    assert set(O.resolve_redrive_from_step_set(st, "fetch-data")) == {"fetch-data", "use"}
    # groups do not leak: a finally step that needs a main step is not part of the main step's closure
    st = O.Story([O.Step("m", ref=True), O.Step("n", ["m"], ref=True)], [], [O.Step("f", ["m"], ref=True), O.Step("g", ["f"], ref=True)])
    assert set(O.resolve_redrive_from_step_set(st, "m")) == {"m", "n"}
    assert set(O.resolve_redrive_from_step_set(st, "f")) == {"f", "g"}


def _stories(seed, count):
    rng = random.Random(seed)
    return [_rand_story(rng)[0] for _ in range(count)]


@pytest.mark.parametrize("seed", range(10))
def test_packed_closure_equals_object_level(seed):
    for story in _stories(4000 + seed, 25):
        ps = P.pack_story(story)
        for name in ps.names:
            want = O.resolve_redrive_from_step_set(story, name)
            mask = closure_packed(ps, ps.index[name])
            got = {ps.names[i] for i in range(ps.S) if (int(mask[i >> 5]) >> (i & 31)) & 1}
            assert got == set(want), (name, got, set(want))


@pytest.mark.gpu
def test_device_closure_equals_oracle():
    from bobrapet_b200 import Frontier
    from bobrapet_b200 import _abi as A
    fr = Frontier(0)
    try:
        stories = _stories(77, 120)
        packed = [P.pack_story(s) for s in stories]
        slots = fr.put_topologies(P.topology_set(packed))
        q_slot, q_step, want = [], [], []
        W = max((p.S + 31) // 32 for p in packed)
        for t, (story, ps) in enumerate(zip(stories, packed)):
            for name in ps.names:
                q_slot.append(int(slots[t]))
                q_step.append(ps.index[name])
                m = np.zeros(W, dtype=np.uint32)
                for n2 in O.resolve_redrive_from_step_set(story, name):
                    i = ps.index[n2]
                    m[i >> 5] |= np.uint32(1 << (i & 31))
                want.append(m)
        got = fr.closure(q_slot, q_step, W)
        assert np.array_equal(got, np.stack(want))
        with pytest.raises(A.FrontierError):
            fr.closure([int(slots[0])], [packed[0].S], W)      # step index out of range
        # a wide random DAG (S = 1000): against the packed statement
        from tests import randgen
        rng = np.random.default_rng(3)
        ts = randgen.random_topologies(rng, 6, 900, 1000)
        sl = fr.put_topologies(ts)
        S_off = np.concatenate(([0], np.cumsum(ts.S.astype(np.int64))))
        R_off = np.concatenate(([0], np.cumsum(ts.S.astype(np.int64) + 1)))
        E_off = np.concatenate(([0], np.cumsum(ts.E.astype(np.int64))))
        for t in range(ts.count):
            S = int(ts.S[t])
            ps = P.PackedStory(["s%d" % i for i in range(S)], {}, ts.row_ptr[R_off[t]:R_off[t] + S + 1],
                               ts.col_idx[E_off[t]:E_off[t + 1]], ts.step_flags[S_off[t]:S_off[t] + S])
            starts = rng.integers(0, S, size=12)
            got = fr.closure([int(sl[t])] * 12, starts, 32)
            for k, st in enumerate(starts):
                assert np.array_equal(got[k, :(S + 31) // 32], closure_packed(ps, int(st))), (t, st)
    finally:
        fr.close()
