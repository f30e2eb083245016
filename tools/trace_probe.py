"""Timeline of the packed-lanes kernel inside a CUDA graph of 8 passes (needs a lib_ab build with -DPACK_TRACE, BF_LIB=...):
per launch and CTA {start, first group landed, last group done, end} from %globaltimer.  Prints, per launch, the gap to the
previous launch's last CTA exit, the spread of CTA starts, the time until the first group is in shared memory, and the tail
(from the first CTA that ran out of groups to the last CTA exit)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bobrapet_b200 import _abi as A, Frontier, synth  # noqa: E402
from bobrapet_b200.records import make_layout  # noqa: E402

cfg = int(os.environ.get("CFG", "3"))
n = int(os.environ.get("RUNS", "100000"))
S = 256
PIPE = os.environ.get("PIPE", "0") == "1"
fr = Frontier(0)
dev = torch.device("cuda", 0)
sets = []
for k in range(3):
    ts = synth.topologies(cfg, k * 10_000_019, n, S)
    slots = fr.put_topologies(ts)
    L = make_layout(S, 0, (A.F_COND | A.F_DECISION) if cfg == 4 else 0)
    st = synth.state(cfg, k * 10_000_019, n, L, slots, ts)
    sets.append((torch.from_numpy(st).to(dev), torch.zeros((n, L.result_stride), dtype=torch.uint8, device=dev),
                 torch.zeros(4, dtype=torch.int64, device=dev)))
ws = torch.cuda.Stream()


def one(k):
    d_state, d_result, d_counts = sets[k % 3]
    if PIPE:
        fr.eval_device(L, n, d_state.data_ptr(), d_result.data_ptr(), d_counts.data_ptr(), ws.cuda_stream, flags=A.EVAL_COUNTS_SET | A.EVAL_PIPELINED)
        return
    d_counts.zero_()
    fr.eval_device(L, n, d_state.data_ptr(), d_result.data_ptr(), d_counts.data_ptr(), ws.cuda_stream)


with torch.cuda.stream(ws):
    for k in range(6):
        one(k)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, stream=ws, capture_error_mode="thread_local"):
    for k in range(8):
        one(k)
with torch.cuda.stream(ws):
    gr.replay()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ws)
    gr.replay()
    e1.record(ws)
torch.cuda.synchronize()
print("graph of 8 passes: %.2f us / pass" % (e0.elapsed_time(e1) * 1e3 / 8))
lib = fr._lib
out = np.zeros((8, 160, 4), dtype=np.uint64)
launch = C.c_uint(0)
rc = lib.bf_debug_pack_trace(out.ctypes.data_as(C.c_void_p), C.byref(launch))
assert rc == 0
G = int(fr.stats()["last_grid"])
tr = out[:, :G, :].astype(np.int64)
order = np.argsort(tr[:, :, 0].min(axis=1))
prev_end = None
for li in order:
    t = tr[li]
    s0 = t[:, 0].min()
    line = "launch %d: CTA starts spread %.2f us | first group landed +%.2f (median CTA, from its own start) | last group done: first CTA +%.2f, last CTA +%.2f | kernel %.2f us" % (
        li, (t[:, 0].max() - s0) / 1e3, np.median(t[:, 1] - t[:, 0]) / 1e3, (t[:, 2].min() - s0) / 1e3, (t[:, 2].max() - s0) / 1e3,
        (t[:, 3].max() - s0) / 1e3)
    if prev_end is not None:
        line += " | gap after previous kernel %.2f us" % ((s0 - prev_end) / 1e3)
    prev_end = t[:, 3].max()
    print(line)
fr.close()
