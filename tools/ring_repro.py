"""tools/ring_repro.py [n]: one cfg3 pass of n runs through bf_eval, compared with the oracle (ring / race debugging)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from bobrapet_b200 import Frontier, synth
from bobrapet_b200.records import make_layout
from oracle import packed as PK
n = int(sys.argv[1]) if len(sys.argv) > 1 else 70001
f = Frontier(0)
synth.set_threads(16)
ts = synth.topologies(3, 0, n, 256)
slots = f.put_topologies(ts)
L = make_layout(256, 0, 0)
state = synth.state(3, 0, n, L, slots, ts)
want, wc = PK.evaluate(PK.PackedTopologies(ts, slots), L, state, threads=16)
worst = 0
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    got, gc = f.eval(L, state)
    bad = np.nonzero((got != want).any(axis=1))[0]
    worst = max(worst, bad.size)
    print("rep", rep, "bad runs", bad.size, "first", bad[:6], "groups t:", sorted(set((bad // 4 // 148).tolist()))[:8], gc == wc, f.stats()["last_block"], f.stats()["last_stages"], flush=True)
sys.exit(1 if worst else 0)
