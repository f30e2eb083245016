"""numpy packers for the record layout of include/bobrafrontier.h.

Host-side helpers used by tests and bench.py to build state records and to take
result records apart.  Pure byte shuffling — no frontier semantics live here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _abi as A


def make_layout(steps_max: int, child_nibbles: int = 0, fields: int = 0) -> A.Layout:
    L = A.Layout()
    rc = A.load().bf_layout_init(C.byref(L), steps_max, child_nibbles, fields)
    if rc != A.BF_OK:
        raise A.FrontierError(rc, "bf_layout_init(steps_max=%d)" % steps_max)
    return L


def layout_py(steps_max: int, child_nibbles: int = 0, fields: int = 0) -> A.Layout:
    """The rule of bf_layout_init restated in Python, for callers that must not load the CUDA library
    (bench.py's CPU reference arm); tests/test_abi_symbols.py checks it field by field against bf_layout_init."""
    def up(v, a):
        return (v + a - 1) // a * a
    L = A.Layout()
    W = (steps_max + 31) // 32
    L.steps_max, L.words, L.fields = steps_max, W, fields
    L.child_nibbles = child_nibbles if fields & A.F_CHILD else 0
    off = 16
    L.off_phase = off
    off += W * 16
    L.off_cond = L.off_decision = L.off_child = A.OFF_NONE
    if fields & A.F_COND:
        L.off_cond = off
        off += W * 8
    if fields & A.F_DECISION:
        L.off_decision = off
        off += W * 8
    if fields & A.F_CHILD:
        L.off_child = off
        off += up((L.child_nibbles + 1) // 2, 4)
    L.state_stride = up(off, 16)
    off = 16
    L.off_ready = off
    off += W * 4
    L.off_skip = off
    off += W * 4
    L.off_fail = L.off_needs_cond = L.off_skip_dep = L.off_phase_out = A.OFF_NONE
    for flag, name, size in ((A.F_OUT_FAIL, "off_fail", 4), (A.F_OUT_NEEDS_COND, "off_needs_cond", 4),
                             (A.F_OUT_SKIP_DEP, "off_skip_dep", 4), (A.F_OUT_PHASE, "off_phase_out", 16)):
        if fields & flag:
            setattr(L, name, off)
            off += W * size
    L.result_stride = up(off, 16)
    return L


def _pack_nibbles(codes: np.ndarray, n_bytes: int) -> np.ndarray:
    """codes [N, C] (values 0..15) -> [N, n_bytes] with item i at byte i/2, low nibble first (child area)."""
    n, s = codes.shape
    buf = np.zeros((n, n_bytes * 2), dtype=np.uint8)
    buf[:, :s] = codes
    return (buf[:, 0::2] | (buf[:, 1::2] << 4)).astype(np.uint8)


def _pack_planes(codes: np.ndarray, words: int, nbits: int) -> np.ndarray:
    """codes [N, S] -> [N, nbits*words*4] uint8: plane b (bit b of every code) = `words` little-endian u32,
    step i at bit i%32 of word i/32."""
    n, s = codes.shape
    buf = np.zeros((n, words * 32), dtype=np.uint8)
    buf[:, :s] = codes
    planes = [np.packbits((buf >> b) & 1, axis=1, bitorder="little") for b in range(nbits)]
    return np.concatenate(planes, axis=1)


def _unpack_planes(raw: np.ndarray, words: int, nbits: int, s: int) -> np.ndarray:
    """inverse of _pack_planes -> codes [N, s]"""
    n = raw.shape[0]
    codes = np.zeros((n, words * 32), dtype=np.uint8)
    for b in range(nbits):
        bits = np.unpackbits(np.ascontiguousarray(raw[:, b * words * 4:(b + 1) * words * 4]), axis=1, bitorder="little")
        codes |= (bits << b).astype(np.uint8)
    return codes[:, :s]


def pack_state(L: A.Layout, slots: np.ndarray, run_flags: np.ndarray, phase: np.ndarray,
               cond: Optional[np.ndarray] = None, decision: Optional[np.ndarray] = None,
               child: Optional[np.ndarray] = None, registered: Optional[np.ndarray] = None,
               out: Optional[np.ndarray] = None) -> np.ndarray:
    """Build [N, state_stride] uint8 state records.

    phase/cond/decision are [N, S] code arrays (S <= steps_max), stored bit-sliced; child is [N, child_nibbles]
    phase codes laid out by bf_topology_child_first; registered is [N] uint64."""
    n = int(slots.shape[0])
    rec = out if out is not None else np.zeros((n, L.state_stride), dtype=np.uint8)
    rec[:, 0:4] = np.ascontiguousarray(slots, dtype="<u4").view(np.uint8).reshape(n, 4)
    rec[:, 4] = run_flags.astype(np.uint8)
    if registered is not None:
        rec[:, 8:16] = np.ascontiguousarray(registered, dtype="<u8").view(np.uint8).reshape(n, 8)
    W = L.words
    rec[:, L.off_phase:L.off_phase + W * 16] = _pack_planes(phase, W, 4)
    if L.off_cond != A.OFF_NONE:
        if cond is None:
            cond = np.zeros_like(phase)
        rec[:, L.off_cond:L.off_cond + W * 8] = _pack_planes(cond, W, 2)
    if L.off_decision != A.OFF_NONE:
        if decision is None:
            decision = np.zeros_like(phase)
        rec[:, L.off_decision:L.off_decision + W * 8] = _pack_planes(decision, W, 2)
    if L.off_child != A.OFF_NONE and child is not None and L.child_nibbles:
        nb = (L.child_nibbles + 1) // 2
        rec[:, L.off_child:L.off_child + nb] = _pack_nibbles(child, nb)
    return rec


def _unpack_bits(words_u8: np.ndarray, s: int) -> np.ndarray:
    """[N, W*4] uint8 (little-endian u32 words) -> [N, s] bool."""
    return np.unpackbits(words_u8, axis=1, bitorder="little")[:, :s].astype(bool)


def unpack_result(L: A.Layout, result: np.ndarray, s: int) -> Dict[str, np.ndarray]:
    n = result.shape[0]
    hdr = np.ascontiguousarray(result[:, 0:16]).view("<u4").reshape(n, 4)
    W = L.words
    out = {"summary": hdr[:, 0].copy(), "n_ready": hdr[:, 1].copy(), "n_skip": hdr[:, 2].copy(),
           "n_expansion": hdr[:, 3].copy()}
    for name, off in (("ready", L.off_ready), ("skip", L.off_skip), ("fail", L.off_fail),
                      ("needs_cond", L.off_needs_cond), ("skip_dep", L.off_skip_dep)):
        if off != A.OFF_NONE:
            out[name] = _unpack_bits(np.ascontiguousarray(result[:, off:off + W * 4]), s)
    if L.off_phase_out != A.OFF_NONE:
        out["phase_out"] = _unpack_planes(result[:, L.off_phase_out:L.off_phase_out + W * 16], W, 4, s)
    return out


TOPO_DTYPE = np.dtype([
    ("n_steps", "<u4"), ("n_edges", "<u4"), ("row_ptr", "<u8"), ("col_idx", "<u8"), ("step_flags", "<u8"),
    ("parallel", "<u8"), ("n_parallel", "<u4"), ("_pad0", "<u4"), ("branch_allow_bits", "<u8"),
    ("n_branch_allow_bits", "<u4"), ("_pad1", "<u4"),
])
assert TOPO_DTYPE.itemsize == C.sizeof(A.Topology), (TOPO_DTYPE.itemsize, C.sizeof(A.Topology))
PAR_DTYPE = np.dtype([("step", "<u2"), ("branches", "<u2"), ("allow_first", "<u4")])
EXP_DTYPE = np.dtype([("run", "<u4"), ("step", "<u2"), ("branch", "<u2")])
