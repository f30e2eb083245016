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


def test_eval_flag_constants_match_the_header_and_the_go_binding():
    """the BF_EVAL_* bits of include/bobrafrontier.h == bobrapet_b200/_abi.py == go/frontier/frontier.go, and they are distinct"""
    text = open(os.path.join(ROOT, "include", "bobrafrontier.h")).read()
    hdr = {n: int(v, 16) for n, v in re.findall(r"#define (BF_EVAL_[A-Z_]+) (0x[0-9a-fA-F]+)u", text)}
    want = {"BF_EVAL_VALIDATE": A.EVAL_VALIDATE, "BF_EVAL_FIXPOINT": A.EVAL_FIXPOINT, "BF_EVAL_EXPANSION": A.EVAL_EXPANSION,
            "BF_EVAL_NO_COUNTS": A.EVAL_NO_COUNTS, "BF_EVAL_CHANGED_ONLY": A.EVAL_CHANGED_ONLY, "BF_EVAL_COUNTS_SET": A.EVAL_COUNTS_SET,
            "BF_EVAL_PIPELINED": A.EVAL_PIPELINED}
    assert hdr == want
    assert len(set(hdr.values())) == len(hdr) and all(v < 0x10000 and v & (v - 1) == 0 for v in hdr.values())   # public bits, one each
    go = open(os.path.join(ROOT, "go", "frontier", "frontier.go")).read()
    for name, go_name in (("BF_EVAL_CHANGED_ONLY", "EvalChangedOnly"), ("BF_EVAL_COUNTS_SET", "EvalCountsSet"), ("BF_EVAL_PIPELINED", "EvalPipelined")):
        m = re.search(r"\b%s\s*=\s*(0x[0-9a-fA-F]+)" % go_name, go)
        assert m and int(m.group(1), 16) == hdr[name], go_name
