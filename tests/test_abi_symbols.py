"""The C-ABI library loads without a GPU and exports every entry point include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

from bobrapet_b200 import _abi as A
from bobrapet_b200.records import make_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bfh?_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = A.load()
    names = _declared("bobrafrontier.h") + _declared("bobrafrontier_host.h")
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), "include/*.h declares %s but the library does not export it" % n
    bound = {n for n, _, _ in A.SYMBOLS}
    assert set(_declared("bobrafrontier.h")) == bound


def test_version_strerror_layout_without_gpu():
    lib = A.load()
    assert lib.bf_abi_version() == A.BF_ABI_VERSION
    assert lib.bf_strerror(A.BF_ETOPO) == b"topology rejected"
    L = make_layout(256, 0, 0)
    assert (L.words, L.state_stride, L.result_stride, L.off_phase, L.off_ready, L.off_skip) == (8, 144, 80, 16, 16, 48)
    L = make_layout(1024, 1024, A.F_COND | A.F_DECISION | A.F_CHILD | A.F_ALL_OUT)
    assert L.state_stride % 16 == 0 and L.result_stride % 16 == 0 and L.off_child == 16 + 512 + 256 + 256
    assert C.sizeof(A.Batch) == 120 and C.sizeof(A.Layout) == 64


def test_struct_sizes_match_header_comments():
    # bf_run_header / bf_result_header are 16 bytes (records start with them)
    assert make_layout(32, 0, 0).off_phase == 16 and make_layout(32, 0, 0).off_ready == 16


def test_python_layout_rule_matches_bf_layout_init():
    """records.layout_py (used by bench.py's CPU arm, which must not load the CUDA library) == bf_layout_init"""
    import itertools
    from bobrapet_b200.records import layout_py
    import bench
    assert (bench.F_COND, bench.F_DECISION, bench.F_CHILD) == (A.F_COND, A.F_DECISION, A.F_CHILD)
    for S, child, fields in itertools.product((1, 31, 32, 33, 64, 256, 1000, 1024), (0, 8, 120, 1024),
                                              (0, A.F_COND, A.F_DECISION | A.F_CHILD, A.F_COND | A.F_DECISION | A.F_CHILD | A.F_ALL_OUT,
                                               A.F_OUT_FAIL | A.F_OUT_PHASE, A.F_CHILD | A.F_OUT_SKIP_DEP | A.F_OUT_NEEDS_COND)):
        assert layout_py(S, child, fields).as_dict() == make_layout(S, child, fields).as_dict(), (S, child, fields)
