"""Seeded synthetic Stories / StoryRuns for the BASELINE.json configurations (data only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .. import _abi as A
from ..frontier import TopologySet
from ..records import PAR_DTYPE

_LIB = None
_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libsynth.so")


def use_library(path: str) -> None:
    """Load the generator from `path` instead (bench.py's CPU arm uses the copy built under oracle/_build so that
    process maps nothing from bobrapet_b200/lib)."""
    global _PATH, _LIB
    _PATH, _LIB = path, None


def set_threads(n: int) -> None:
    """Runs are independent PRNG streams: generate them on `n` OpenMP threads."""
    _lib().synth_set_threads(int(n))

# (S, parallel steps P, branches B) of the named configurations
CONFIGS = {1: (3, 0, 0), 2: (64, 0, 0), 3: (256, 0, 0), 4: (256, 0, 0), 5: (1024, 8, 128)}


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            raise RuntimeError("%s missing: run __graft_entry__.build()" % _PATH)
        lib = C.CDLL(_PATH)
        lib.synth_edges.restype = C.c_uint32
        lib.synth_edges.argtypes = [C.c_uint32, C.c_uint32]
        lib.synth_topologies.restype = None
        lib.synth_topologies.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
        lib.synth_state.restype = None
        lib.synth_state.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(A.Layout), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.synth_set_threads.restype = None
        lib.synth_set_threads.argtypes = [C.c_int]
        _LIB = lib
    return _LIB


def topologies(cfg: int, run_lo: int, n: int, S: int = 0) -> TopologySet:
    """Topologies of runs [run_lo, run_lo+n), one per run ("unique-topology mode")."""
    lib = _lib()
    S = S or CONFIGS[cfg][0]
    E = lib.synth_edges(cfg, S)
    P, B = (S // 128, 128) if cfg == 5 else (0, 0)
    row_ptr = np.zeros(n * (S + 1), dtype=np.uint32)
    col = np.zeros(n * E, dtype=np.uint16)
    flags = np.zeros(n * S, dtype=np.uint8)
    par = np.zeros(n * P, dtype=PAR_DTYPE)
    allow = np.zeros(n * P * B // 8, dtype=np.uint8)
    lib.synth_topologies(cfg, run_lo, n, S, row_ptr.ctypes.data, col.ctypes.data, flags.ctypes.data,
                         par.ctypes.data if P else None, allow.ctypes.data if P else None)
    return TopologySet(np.full(n, S, np.uint32), np.full(n, E, np.uint32), row_ptr, col, flags,
                       np.full(n, P, np.uint32), par, allow)


def state(cfg: int, run_lo: int, n: int, L: A.Layout, slots: np.ndarray, ts: TopologySet,
          child_first: np.ndarray = None, out: np.ndarray = None) -> np.ndarray:
    """State records [n, state_stride] for runs [run_lo, run_lo+n) whose topologies are `ts`."""
    lib = _lib()
    S = int(ts.S[0])
    P = int(ts.P[0]) if ts.P.size else 0
    B = int(ts.parallel["branches"][0]) if P else 0
    rec = out if out is not None else np.zeros((n, L.state_stride), dtype=np.uint8)
    assert rec.flags["C_CONTIGUOUS"] and rec.shape == (n, L.state_stride)
    slots = np.ascontiguousarray(slots, dtype=np.uint32)
    cf = np.ascontiguousarray(child_first if child_first is not None else np.zeros(max(P, 1)), dtype=np.uint32)
    lib.synth_state(cfg, run_lo, n, S, C.byref(L), slots.ctypes.data, ts.step_flags.ctypes.data, cf.ctypes.data, P, B,
                    rec.ctypes.data)
    return rec


def shared_topology_slots(n_runs: int, n_topologies: int) -> np.ndarray:
    """Shared-topology mode: run r uses topology r % D."""
    return (np.arange(n_runs, dtype=np.uint64) % np.uint64(n_topologies)).astype(np.uint32)
