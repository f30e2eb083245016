#!/bin/bash
# tools/sanitize.sh: compute-sanitizer (memcheck, racecheck, synccheck) over small passes of every kernel shape — smoke() plus a
# wide (S = 1024, parallel joins, 10-bit rows) and a pipelined batch.  Prints the tools' summaries.
cd "$(dirname "$0")/.."
cat > /tmp/san_probe.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import __graft_entry__ as g
from bobrapet_b200 import _abi as A, Frontier, synth
from bobrapet_b200.records import make_layout
from oracle import packed as PK
g.smoke()
fr = Frontier(0)
ts = synth.topologies(5, 0, 96, 1024)
slots = fr.put_topologies(ts)
pt = PK.PackedTopologies(ts, slots)
child = pt.max_child_nibbles()
L = make_layout(1024, child, A.F_COND | A.F_DECISION | A.F_CHILD)
cf = fr.child_first(int(slots[0]))
st = synth.state(5, 0, 96, L, slots, ts, cf)
got, c = fr.eval(L, st)
want, wc = PK.evaluate(pt, L, st, threads=4)
assert np.array_equal(got, want) and c == wc
f2 = Frontier(0)
ts3 = synth.topologies(3, 0, 1500, 256)
s3 = f2.put_topologies(ts3)
L3 = make_layout(256, 0, 0)
st3 = synth.state(3, 0, 1500, L3, s3, ts3)
w3, wc3 = PK.evaluate(PK.PackedTopologies(ts3, s3), L3, st3, threads=4)
dev = torch.device("cuda", 0)
d_state = torch.from_numpy(st3).to(dev)
res = [torch.zeros((1500, L3.result_stride), dtype=torch.uint8, device=dev) for _ in range(3)]
cnt = [torch.full((4,), 9, dtype=torch.int64, device=dev) for _ in range(3)]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for k in range(6):
        f2.eval_device(L3, 1500, d_state.data_ptr(), res[k % 3].data_ptr(), cnt[k % 3].data_ptr(), s.cuda_stream,
                       flags=A.EVAL_COUNTS_SET | A.EVAL_PIPELINED)
torch.cuda.synchronize()
for k in range(3):
    assert np.array_equal(res[k].cpu().numpy(), w3) and cnt[k].cpu().numpy().tolist() == [wc3["ready"], wc3["skip"], wc3["expansion"], wc3["evals"]]
print("probe ok")
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_probe.py 2>&1 | grep -v "^smoke\|NCCL" | tail -6
done
