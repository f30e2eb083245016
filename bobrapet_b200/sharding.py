"""Multi-GPU sharding of a StoryRun batch.

StoryRuns are independent (no step of one run reads another run), so a batch shards by
contiguous blocks of ceil(N/G) runs per rank with NO data-path collective.  The one exchange
the path has is an all-gather of each shard's counts (ready, skip, expansion, evals — 32
bytes per rank), which gives every rank the global offsets of its compacted ready /
expansion lists.  Works with torch.distributed NCCL (CUDA tensors, one process per GPU) and
gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist

COUNT_FIELDS = ("ready", "skip", "expansion", "evals")


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of ceil(N/G) runs for `rank` (the last ranks may be short or empty)."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


class CountExchange:
    """All-gather of the per-shard bf_counts block (4 x int64)."""

    def __init__(self, device: torch.device, world: int):
        self.world = world
        self.device = device
        self.comm_stream = torch.cuda.Stream(device) if device.type == "cuda" and world > 1 else None

    def new_buffer(self) -> torch.Tensor:
        return torch.zeros(4 * self.world, dtype=torch.int64, device=self.device)

    def gather(self, counts: torch.Tensor, out: torch.Tensor, compute_stream=None) -> None:
        """Enqueue the all-gather of `counts` into `out`; on CUDA it runs on a side stream ordered after
        `compute_stream`, so it overlaps the next pass."""
        if self.world == 1:
            out.copy_(counts)
            return
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(compute_stream)
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_gather_into_tensor(out, counts)
        else:
            dist.all_gather_into_tensor(out, counts)

    def join(self, compute_stream=None) -> None:
        if self.comm_stream is not None and compute_stream is not None:
            compute_stream.wait_stream(self.comm_stream)


def global_offsets(gathered: torch.Tensor, rank: int) -> Dict[str, Dict[str, int]]:
    """From the gathered [world*4] counts: totals and this rank's exclusive offsets per list."""
    g = gathered.view(-1, 4).cpu()
    excl = torch.cumsum(g, dim=0) - g
    return {
        "total": {f: int(g[:, i].sum()) for i, f in enumerate(COUNT_FIELDS)},
        "offset": {f: int(excl[rank, i]) for i, f in enumerate(COUNT_FIELDS)},
        "mine": {f: int(g[rank, i]) for i, f in enumerate(COUNT_FIELDS)},
    }
