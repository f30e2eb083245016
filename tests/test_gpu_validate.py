"""Row f3: batched graph validation on the device vs the reference's Kahn (oracle/pyoracle.py
validate_runtime_dependency_graph, dag.go:3076-3146) — cycles rejected, level counts exact."""
import numpy as np
import pytest

from bobrapet_b200 import _abi as A
from bobrapet_b200 import Frontier
from bobrapet_b200.frontier import TopologySet
from oracle import pyoracle as O
from tests import randgen

pytestmark = pytest.mark.gpu


def _levels_and_cycle(S, rp, ci):
    """(has_cycle, levels) by the definition the kernel uses; cross-checked against the oracle's Kahn below."""
    done = np.zeros(S, bool)
    levels = 0
    while True:
        newly = [i for i in range(S) if not done[i] and all(done[d] for d in ci[rp[i]:rp[i + 1]])]
        if not newly:
            break
        done[newly] = True
        levels += 1
    return (not done.all()), levels


def test_device_validation_matches_kahn():
    rng = np.random.default_rng(42)
    ts = randgen.random_topologies(rng, 120, 1, 200, parallel=False)
    # inject back-edges into every third topology: i needs j where j (transitively) needs i, or a self-loop
    rp_all, ci_all, E = [], [], []
    pos_rp = pos_ci = 0
    expect = []
    for t in range(ts.count):
        S, e = int(ts.S[t]), int(ts.E[t])
        rp = ts.row_ptr[pos_rp:pos_rp + S + 1].astype(np.int64).copy()
        ci = ts.col_idx[pos_ci:pos_ci + e].astype(np.int64).copy()
        pos_rp += S + 1
        pos_ci += e
        rows = [list(ci[rp[i]:rp[i + 1]]) for i in range(S)]
        if t % 3 == 0 and S >= 2:
            if t % 2 == 0:
                i = int(rng.integers(0, S))
                rows[i].append(i)                       # self-loop (dag_test.go:321)
            else:
                cand = [i for i in range(S) if rows[i]]
                if cand:
                    i = int(rng.choice(cand)); d = int(rows[i][0])
                    rows[d].append(i)                   # 2-cycle through an existing edge
        rows = [sorted(set(r)) for r in rows]
        rp2 = np.zeros(S + 1, np.uint32); rp2[1:] = np.cumsum([len(r) for r in rows])
        ci2 = np.asarray([c for r in rows for c in r], np.uint16)
        rp_all.append(rp2); ci_all.append(ci2); E.append(len(ci2))
        cyc, lv = _levels_and_cycle(S, rp2, ci2)
        steps = [O.Step("s%d" % i, needs=["s%d" % d for d in rows[i]]) for i in range(S)]
        err = O.validate_runtime_dependency_graph(steps)
        assert (err is not None) == cyc and (not cyc or "cycle" in err)
        expect.append((cyc, lv))
    bad = TopologySet(ts.S, E, np.concatenate(rp_all), np.concatenate(ci_all), ts.step_flags)
    fr = Frontier(0)
    try:
        with pytest.raises(A.FrontierError) as e:
            fr.put_topologies(bad)                       # host Kahn rejects the batch
        assert "cycle" in str(e.value)
        slots, status = fr.put_topologies_checked_on_device(bad)
        for t, (cyc, lv) in enumerate(expect):
            assert bool(status[t] & 1) == cyc, t
            assert int(status[t] >> 8) == lv, (t, int(status[t] >> 8), lv)
            assert (slots[t] == 0xFFFFFFFF) == cyc
        good = slots[slots != 0xFFFFFFFF]
        again = fr.check_topologies(good)
        assert not (again & 1).any()
        assert fr.check_topologies(np.asarray([0x7FFFFFF0], np.uint32))[0] == 0xFFFFFFFF
    finally:
        fr.close()
