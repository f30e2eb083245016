"""N>1 path on CPU: two gloo ranks shard a batch, evaluate their shards (with the CPU oracle standing in
for the kernel), all-gather the counts and agree with the unsharded result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bobrapet_b200 import _abi as A, synth
from bobrapet_b200.records import make_layout
from bobrapet_b200.sharding import CountExchange, global_offsets, shard_range
from oracle import packed as PK

N_TOTAL, S, CFG = 1001, 96, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _eval_range(lo, hi):
    n = hi - lo
    L = make_layout(S, 0, A.F_COND | A.F_DECISION)
    if n == 0:
        return L, np.zeros((0, L.result_stride), np.uint8), {"ready": 0, "skip": 0, "expansion": 0, "evals": 0}
    ts = synth.topologies(CFG, lo, n, S)
    st = synth.state(CFG, lo, n, L, np.arange(n, dtype=np.uint32), ts)
    res, counts = PK.evaluate(PK.PackedTopologies(ts), L, st)
    return L, res, counts


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N_TOTAL, world, rank)
    _, res, counts = _eval_range(lo, hi)
    ex = CountExchange(torch.device("cpu"), world)
    mine = torch.tensor([counts[f] for f in ("ready", "skip", "expansion", "evals")], dtype=torch.int64)
    out = ex.new_buffer()
    ex.gather(mine, out)
    q.put((rank, lo, hi, global_offsets(out, rank), res.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    for n, w in ((10, 3), (8, 8), (5, 8), (0, 2), (100000, 8)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_two_rank_gloo_matches_unsharded():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, full_res, full_counts = _eval_range(0, N_TOTAL)
    joined = b"".join(g[4] for g in got)
    assert joined == full_res.tobytes()
    for rank, lo, hi, off, _ in got:
        assert off["total"] == full_counts
    assert got[0][3]["offset"]["ready"] == 0
    assert got[1][3]["offset"]["ready"] == got[0][3]["mine"]["ready"]
    assert got[1][3]["offset"]["evals"] == (got[0][2] - got[0][1]) * S
