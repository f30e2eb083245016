"""The device topology records, built on the host without a GPU (bf_topology_record_build) and decoded here by the rules of
bobrapet_b200/csrc/device_record.h: whatever row format a topology gets — CSR, fixed-width rows with u16 / byte / 10-bit entries —
the rows must say exactly the `needs` sets the caller gave (short rows repeat an entry, rows without needs are flagged NODEP and
point at themselves), the static planes must be the step flags, and the size must be the one the format promises."""
import ctypes as C
import struct

import numpy as np
import pytest

from bobrapet_b200 import _abi as A
from tests import randgen

HDR = struct.Struct("<HHHHHHHHHHIII")   # TopoHeader, 32 bytes


def _build(ts, t, monkeypatch=None):
    lib = A.load()
    desc = ts.descriptors()
    n = C.c_uint32()
    d = desc[t:t + 1]
    rc = lib.bf_topology_record_build(C.cast(d.ctypes.data, C.POINTER(A.Topology)), None, 0, C.byref(n))
    assert rc == A.BF_ENOMEM and n.value % 16 == 0
    buf = np.zeros(n.value, dtype=np.uint8)
    assert lib.bf_topology_record_build(C.cast(d.ctypes.data, C.POINTER(A.Topology)), buf.ctypes.data, n.value, C.byref(n)) == A.BF_OK
    return buf


def _decode_rows(rec):
    S, W, max_deg, P, n_main, n_comp, n_final, child_nib, off_col, ell, off_planes, off_par, rec_bytes = HDR.unpack_from(rec.tobytes(), 0)
    assert rec_bytes == rec.shape[0] and W == (S + 31) // 32
    planes = rec[off_planes:].view("<u4")
    bit = lambda pl, i: (int(planes[pl * W + (i >> 5)]) >> (i & 31)) & 1   # noqa: E731
    rows = []
    if ell == 0:
        rp = rec[32:32 + 2 * (S + 1)].view("<u2").astype(np.int64)
        ci = rec[off_col:].view("<u2")
        rows = [sorted(int(x) for x in ci[rp[i]:rp[i + 1]]) for i in range(S)]
        fmt = "csr"
    else:
        K = ell & 0xFF
        col = rec[off_col:]
        for i in range(S):
            if ell & 0x200:
                lo = int(col[4 * i:4 * i + 4].view("<u4")[0])
                e = [lo & 0x3FF, (lo >> 10) & 0x3FF, (lo >> 20) & 0x3FF, (lo >> 30) | (int(col[128 * W + i]) << 2)]
            elif ell & 0x100:
                e = [int(x) for x in col[K * i:K * i + K]]
            else:
                e = [int(x) for x in col[2 * K * i:2 * K * i + 2 * K].view("<u2")]
            if ell & 0x300:
                if bit(8, i):                     # NODEP: the row points at itself
                    assert set(e) == {i}
                    e = []
                else:
                    assert e[0] in e and all(x < S for x in e)
            else:
                e = [x for x in e if x != 32 * W]  # PAD entries
            rows.append(sorted(set(e)))
        fmt = "p10" if ell & 0x200 else ("byte%d" % K if ell & 0x100 else "u16x%d" % K)
        if ell & 0x300:
            for i in range(S, 32 * W):
                assert bit(8, i)                  # the rows past S are NODEP as well
    flags = np.zeros(S, np.uint8)
    for i in range(S):
        t = bit(0, i) | (bit(1, i) << 1) | (bit(2, i) << 2)
        g = 1 if bit(6, i) else (2 if bit(7, i) else 0)
        flags[i] = t | (bit(3, i) * A.SF_ALLOW_FAILURE) | (bit(4, i) * A.SF_ON_TIMEOUT_SKIP) | (bit(5, i) * A.SF_HAS_IF) | (g << A.SF_GROUP_SHIFT)
    return fmt, rows, flags, (S, W, max_deg, P, n_main, n_comp, n_final)


@pytest.mark.parametrize("forced", [None, "ell16", "csr"])
@pytest.mark.parametrize("smin,smax,deg,fill", [(1, 40, 4, 0.9), (200, 256, 4, 0.9), (225, 256, 2, 0.8), (257, 512, 4, 0.9),
                                               (513, 1024, 4, 0.9), (1024, 1024, 3, 0.6), (600, 900, 2, 0.9), (900, 1024, 5, 0.3)])
def test_records_say_what_the_caller_gave(monkeypatch, forced, smin, smax, deg, fill):
    if forced:
        monkeypatch.setenv("BF_TOPO_FORMAT", forced)
    else:
        monkeypatch.delenv("BF_TOPO_FORMAT", raising=False)
    rng = np.random.default_rng(smin * 7 + smax + deg)
    ts = randgen.random_topologies(rng, 6, smin, smax, max_deg=deg, parallel=False, fill=fill)
    R_off = np.concatenate(([0], np.cumsum(ts.S.astype(np.int64) + 1)))
    E_off = np.concatenate(([0], np.cumsum(ts.E.astype(np.int64))))
    S_off = np.concatenate(([0], np.cumsum(ts.S.astype(np.int64))))
    seen = set()
    for t in range(ts.count):
        S = int(ts.S[t])
        rec = _build(ts, t)
        fmt, rows, flags, (hS, W, max_deg, P, n_main, n_comp, n_final) = _decode_rows(rec)
        seen.add(fmt)
        rp = ts.row_ptr[R_off[t]:R_off[t] + S + 1].astype(np.int64)
        ci = ts.col_idx[E_off[t]:E_off[t + 1]].astype(np.int64)
        want = [sorted(int(x) for x in ci[rp[i]:rp[i + 1]]) for i in range(S)]
        assert hS == S and rows == want, (fmt, t)
        sf = ts.step_flags[S_off[t]:S_off[t] + S]
        groups = (sf & 0xC0) >> A.SF_GROUP_SHIFT
        assert np.array_equal(flags & 0x3F, sf & 0x3F) and np.array_equal((flags >> A.SF_GROUP_SHIFT) & 3, np.minimum(groups, 2))
        assert (n_main, n_comp, n_final) == (int((groups == 0).sum()), int((groups == 1).sum()), int((groups >= 2).sum()))
        md = int((rp[1:] - rp[:-1]).max()) if S else 0
        assert max_deg == md
        # the format each topology must get (plan_record): byte entries up to 256 steps, 10-bit rows of four above 512, u16 rows
        # in between and for rows of two, CSR for longer rows or when forced
        if forced == "csr" or md > 4:
            assert fmt == "csr"
        elif forced is None and 32 * W <= 256:
            assert fmt in ("byte2", "byte4", "csr")
        elif forced is None and W > 16 and md > 2:
            assert fmt in ("p10", "csr")
        elif 32 * W == 256:
            assert fmt == "csr"               # never u16 rows at exactly 8 words
        else:
            assert fmt in ("u16x2", "u16x4", "csr")
    assert seen


def test_rejections_without_a_gpu():
    lib = A.load()
    rng = np.random.default_rng(5)
    ts = randgen.random_topologies(rng, 1, 30, 30, max_deg=3, parallel=False, forward_refs=False)
    ts.col_idx[0] = 40                                     # unknown step dependency (dag_test.go:206)
    n = C.c_uint32()
    d = ts.descriptors()
    assert lib.bf_topology_record_build(C.cast(d.ctypes.data, C.POINTER(A.Topology)), None, 0, C.byref(n)) == A.BF_ETOPO
    S = 6                                                  # a 3-cycle 1 -> 2 -> 3 -> 1
    rows = [[], [3], [1], [2], [], []]
    rp = np.zeros(S + 1, np.uint32); rp[1:] = np.cumsum([len(r) for r in rows])
    from bobrapet_b200.frontier import TopologySet
    cyc = TopologySet([S], [3], rp, np.asarray([c for r in rows for c in r], np.uint16), np.zeros(S, np.uint8))
    d = cyc.descriptors()
    assert lib.bf_topology_record_build(C.cast(d.ctypes.data, C.POINTER(A.Topology)), None, 0, C.byref(n)) == A.BF_ETOPO
