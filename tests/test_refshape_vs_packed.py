"""Third implementation: the reference-shaped C++ restatement (string-keyed maps, graph rebuilt per pass)
against the packed C oracle on adversarial random batches and the synthetic BASELINE configurations."""
import numpy as np
import pytest

from bobrapet_b200 import _abi as A, synth
from bobrapet_b200.records import make_layout, unpack_result
from oracle import packed as PK
from tests import randgen

ALL = A.F_COND | A.F_DECISION | A.F_ALL_OUT


def _check(pt, L, state, S):
    want, wc = PK.evaluate(pt, L, state, 0, 0, 4)
    rs = PK.RefShapeBatch(pt, L, state)
    got, evals = rs.run(threads=4)
    rs.close()
    assert evals == wc["evals"]
    if not np.array_equal(got, want):
        bad = np.nonzero((got != want).any(axis=1))[0]
        r = int(bad[0])
        g, w = unpack_result(L, got[r:r + 1], S), unpack_result(L, want[r:r + 1], S)
        diff = {k: (g[k][0], w[k][0]) for k in g if not np.array_equal(g[k], w[k])}
        raise AssertionError("%d runs differ; first %d: %s" % (len(bad), r, diff))


@pytest.mark.parametrize("seed", range(6))
def test_random_adversarial(seed):
    rng = np.random.default_rng(300 + seed)
    ts = randgen.random_topologies(rng, 25, 1, [20, 70, 33, 130, 64, 9][seed])
    slots = np.arange(ts.count, dtype=np.uint32)
    L, state, _ = randgen.random_state(rng, ts, slots, 600, ALL, phase_mix=("any" if seed % 2 else "progress"))
    _check(PK.PackedTopologies(ts, slots), L, state, int(ts.S.max()))


@pytest.mark.parametrize("cfg,n,S", [(2, 300, 64), (3, 200, 256), (4, 200, 256), (5, 40, 1024)])
def test_synthetic_configs(cfg, n, S):
    ts = synth.topologies(cfg, 0, n, S)
    pt = PK.PackedTopologies(ts)
    child = pt.max_child_nibbles()
    L = make_layout(S, child, ALL | (A.F_CHILD if child else 0))
    state = synth.state(cfg, 0, n, L, np.arange(n, dtype=np.uint32), ts, pt.child_first[:int(ts.P[0])] if child else None)
    _check(pt, L, state, S)
