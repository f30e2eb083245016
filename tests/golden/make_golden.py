"""Generates the committed golden vectors tests/golden/*.npz.

    python tests/golden/make_golden.py

The reference (Go) cannot run in this image, so the vectors come from the oracle AFTER it has been pinned by the
reference's own known-answer tests (tests/test_oracle_kat.py): seeded inputs of the BASELINE.json configurations at
small sizes plus adversarial random batches, and the whole result records / counts / expansion tuples the packed
oracle (oracle/packed_ref.c) produces for them, single pass and fixpoint.  tests/test_golden.py checks the oracle
against these files on CPU (drift detection) and the CUDA path against them on the GPU (no oracle involved there).
State records carry the TOPOLOGY INDEX in the slot field; the test patches in the slots the device assigns.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bobrapet_b200 import _abi as A  # noqa: E402
from bobrapet_b200 import synth  # noqa: E402
from bobrapet_b200.records import make_layout  # noqa: E402
from oracle import packed as PK  # noqa: E402
from tests import randgen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALL = A.F_COND | A.F_DECISION | A.F_ALL_OUT


def dump(name, ts, L_args, state, topo_of_run):
    slots = np.arange(ts.count, dtype=np.uint32)          # slot == topology index in the golden state
    pt = PK.PackedTopologies(ts, slots)
    L = make_layout(*L_args)
    out = {"S": ts.S, "E": ts.E, "P": ts.P, "row_ptr": ts.row_ptr, "col_idx": ts.col_idx, "step_flags": ts.step_flags,
           "par_step": ts.parallel["step"], "par_branches": ts.parallel["branches"], "par_allow_first": ts.parallel["allow_first"],
           "allow_bits": ts.allow_bits, "layout_args": np.asarray(L_args, dtype=np.uint32), "state": state,
           "topo_of_run": np.asarray(topo_of_run, dtype=np.uint32)}
    for tag, flags in (("single", 0), ("fixpoint", A.EVAL_FIXPOINT)):
        res, counts = PK.evaluate(pt, L, state, flags, 0, threads=1)
        exp, n = PK.expand(pt, L, state, res, int(counts["expansion"]) + 1)
        out["result_" + tag] = res
        out["counts_" + tag] = np.asarray([counts["ready"], counts["skip"], counts["expansion"], counts["evals"]], dtype=np.uint64)
        out["expansion_" + tag] = exp[:n]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s %7d bytes  runs %d" % (name + ".npz", os.path.getsize(path), state.shape[0]))


def main():
    for cfg, n in ((2, 48), (3, 32), (4, 32), (5, 6)):
        S = synth.CONFIGS[cfg][0]
        ts = synth.topologies(cfg, 0, n, S)
        slots = np.arange(n, dtype=np.uint32)
        pt = PK.PackedTopologies(ts, slots)
        child = pt.max_child_nibbles()
        fields = (ALL if cfg in (4, 5) else A.F_ALL_OUT) | (A.F_CHILD if child else 0)
        L_args = (S, child, fields)
        L = make_layout(*L_args)
        cf = pt.child_first[:int(ts.P[0])] if child else None
        st = synth.state(cfg, 0, n, L, slots, ts, cf)
        dump("cfg%d_small" % cfg, ts, L_args, st, slots)
    for k, (seed, smax, mix) in enumerate(((11, 70, "any"), (12, 300, "progress"))):
        rng = np.random.default_rng(seed)
        ts = randgen.random_topologies(rng, 12, 1, smax)
        slots = np.arange(ts.count, dtype=np.uint32)
        L, st, topo = randgen.random_state(rng, ts, slots, 64, ALL, phase_mix=mix)
        dump("random%d" % k, ts, (L.steps_max, L.child_nibbles, L.fields), st, topo)


if __name__ == "__main__":
    main()
