"""e2e sweep: bf_eval (pinned host buffers) over BF_E2E_CHUNKS values at BASELINE configs[2].  Prints ms per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bobrapet_b200 import _abi as A, Frontier, synth
from bobrapet_b200.records import make_layout

n, S = 100_000, 256
fr = Frontier(0)
ts = synth.topologies(3, 0, n, S)
slots = fr.put_topologies(ts)
L = make_layout(S, 0, 0)
st = synth.state(3, 0, n, L, slots, ts)
hs = fr.alloc_pinned(n * L.state_stride).reshape(n, L.state_stride)
hr = fr.alloc_pinned(n * L.result_stride).reshape(n, L.result_stride)
hs[:] = st
for ch in [int(x) for x in (sys.argv[1:] or ["1", "2", "4", "8", "12", "16", "24", "32"])]:
    os.environ["BF_E2E_CHUNKS"] = str(ch)
    for _ in range(3):
        fr.eval(L, hs, hr)
    t0 = time.perf_counter()
    K = 40
    for _ in range(K):
        fr.eval(L, hs, hr)
    dt = (time.perf_counter() - t0) / K
    print("chunks %2d  %.3f ms/call  %.3e evals/s  (%d B in, %d B out)" % (ch, dt * 1e3, n * S / dt, hs.nbytes, hr.nbytes), flush=True)
fr.close()
