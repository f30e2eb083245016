"""tools/e2e_probe.py: where a resident tick's time goes (100k runs x 256 steps).  Variants of bf_resident_tick_compact /
bf_resident_tick timed over 30 calls each, pinned host buffers."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bobrapet_b200 import _abi as A, Frontier, synth  # noqa: E402
from bobrapet_b200.records import make_layout  # noqa: E402

n, S = 100000, 256
fr = Frontier(0)
synth.set_threads(16)
ts = synth.topologies(3, 0, n, S)
slots = fr.put_topologies(ts)
L = make_layout(S, 0, 0)
state = synth.state(3, 0, n, L, slots, ts)
h = fr.resident_create(L, n)
fr.resident_upload(h, 0, state)
rng = np.random.default_rng(1)
hr = fr.alloc_pinned(n * L.result_stride).reshape(n, L.result_stride)
h_head = fr.alloc_pinned(n * 4).view(np.uint32)
cap = 600000
h_ev = fr.alloc_pinned(cap * 2).view(np.uint16)


def deltas(rate):
    k = max(1, int(n * S * rate))
    flat = rng.choice(n * S, size=k, replace=False)
    d = fr.alloc_pinned(k * 8).view(fr.DELTA_DTYPE)
    d["run"], d["index"], d["field"] = flat // S, flat % S, A.DELTA_PHASE
    d["code"] = rng.choice([0, 2, 3, 3, 3, 4, 13], size=k)
    return d


def timeit(name, fn, reps=30):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    print("%-58s %7.1f us/call  %.3e evals/s" % (name, 1e6 * dt, n * S / dt), flush=True)


empty = np.zeros(0, dtype=fr.DELTA_DTYPE)
for rate in (0.01, 0.004, 0.001):
    ds = [deltas(rate) for _ in range(3)]
    i = [0]

    def tick():
        i[0] += 1
        return fr.resident_tick_compact(h, n, ds[i[0] % 3], cap, head=h_head, events=h_ev)
    timeit("tick_compact, %.1f%% deltas (%d)" % (100 * rate, ds[0].shape[0]), tick)
ds1 = [deltas(0.01) for _ in range(3)]
j = [0]


def tick_changed():
    j[0] += 1
    return fr.resident_tick_compact(h, n, ds1[j[0] % 3], cap, flags=A.EVAL_CHANGED_ONLY, head=h_head, events=h_ev)


timeit("tick_compact changed-only, 1% deltas", tick_changed)
print("   listed runs / events of the last changed-only tick:", tick_changed()[4], tick_changed()[2])
timeit("tick_compact, no deltas", lambda: fr.resident_tick_compact(h, n, empty, cap, head=h_head, events=h_ev))
timeit("tick_compact, no deltas, events_cap 0", lambda: fr.resident_tick_compact(h, n, empty, 0, head=h_head, events=h_ev))
timeit("tick dense (mask records), no deltas", lambda: fr.resident_tick(h, L, n, empty, hr))
ds = [deltas(0.01) for _ in range(3)]
timeit("tick dense, 1% deltas", lambda: fr.resident_tick(h, L, n, ds[0], hr))
import torch  # noqa: E402
d_state = torch.from_numpy(state).cuda()
d_res = torch.zeros((n, L.result_stride), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def kern():
    fr.eval_device(L, n, d_state.data_ptr(), d_res.data_ptr(), 0, st, flags=A.EVAL_NO_COUNTS)
    torch.cuda.synchronize()
timeit("bf_eval_device + sync (kernel only)", kern)
print(fr.stats())
