"""ctypes mirror of include/bobrafrontier.h and loader of lib/libbobrafrontier.so.

The product path has NO CPU fallback: if the CUDA library is missing or cannot
be loaded this module raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BF_LIB") or os.path.join(_HERE, "lib", "libbobrafrontier.so")  # BF_LIB: A/B builds

BF_ABI_VERSION = 3
BF_OK, BF_EINVAL, BF_ENOMEM, BF_ECUDA, BF_ENCCL, BF_ETOPO, BF_ENODEV = 0, -1, -2, -3, -4, -5, -6

# phase codes (pkg/enums/enums.go:44-97 order; 14 = Pending + "Queued due to ..." message)
PHASE_NONE, PHASE_PENDING, PHASE_RUNNING, PHASE_SUCCEEDED, PHASE_FAILED = 0, 1, 2, 3, 4
PHASE_FINISHED, PHASE_CANCELED, PHASE_COMPENSATED, PHASE_PAUSED, PHASE_BLOCKED = 5, 6, 7, 8, 9
PHASE_SCHEDULING, PHASE_TIMEOUT, PHASE_ABORTED, PHASE_SKIPPED, PHASE_PENDING_QUEUED = 10, 11, 12, 13, 14
PHASE_NAMES = ["", "Pending", "Running", "Succeeded", "Failed", "Finished", "Canceled", "Compensated",
               "Paused", "Blocked", "Scheduling", "Timeout", "Aborted", "Skipped", "Pending"]

QUEUED_NONE, QUEUED_PRIORITY, QUEUED_GLOBAL, QUEUED_QUEUE, QUEUED_OTHER = range(5)  # BF_QUEUED_*
SCHED_NONE = 0xFFFFFFFF
DELTA_PHASE, DELTA_COND, DELTA_DECISION, DELTA_CHILD, DELTA_RUN_FLAGS, DELTA_REGISTERED, DELTA_TOPO_SLOT = range(7)  # BF_DELTA_*
STEP_ENGRAM, STEP_CONDITION, STEP_PARALLEL, STEP_SLEEP, STEP_STOP, STEP_WAIT, STEP_EXECUTE_STORY, STEP_GATE = range(8)
STEP_TYPE_CODE = {"": STEP_ENGRAM, "condition": STEP_CONDITION, "parallel": STEP_PARALLEL, "sleep": STEP_SLEEP,
                  "stop": STEP_STOP, "wait": STEP_WAIT, "executeStory": STEP_EXECUTE_STORY, "gate": STEP_GATE}
SF_TYPE_MASK, SF_ALLOW_FAILURE, SF_ON_TIMEOUT_SKIP, SF_HAS_IF, SF_GROUP_SHIFT = 0x07, 0x08, 0x10, 0x20, 6
GROUP_MAIN, GROUP_COMPENSATION, GROUP_FINALLY, GROUP_DONE = 0, 1, 2, 3
RF_FAIL_FAST, RF_REALTIME, RF_TOPOLOGY_TERMINATED, RF_HOST_GROUP, RF_HOST_GROUP_SHIFT = 0x01, 0x02, 0x04, 0x08, 4
COND_PASS, COND_SKIP, COND_HOLD, COND_FAIL = 0, 1, 2, 3
DEC_PENDING, DEC_SUCCEED, DEC_FAIL, DEC_TIMED_OUT = 0, 1, 2, 3
MAX_STEPS, MAX_EDGES, MAX_PARALLEL, OFF_NONE = 1024, 65535, 64, 0xFFFFFFFF

F_COND, F_DECISION, F_CHILD = 0x1, 0x2, 0x4
F_OUT_FAIL, F_OUT_NEEDS_COND, F_OUT_SKIP_DEP, F_OUT_PHASE = 0x10, 0x20, 0x40, 0x80
F_ALL_OUT = F_OUT_FAIL | F_OUT_NEEDS_COND | F_OUT_SKIP_DEP | F_OUT_PHASE
EVAL_VALIDATE, EVAL_FIXPOINT, EVAL_EXPANSION, EVAL_NO_COUNTS = 0x1, 0x2, 0x4, 0x8

SUM_GROUP_MASK, SUM_MAIN_DONE, SUM_MAIN_FAILED, SUM_COMP_DONE, SUM_FINAL_DONE = 0x3, 0x4, 0x8, 0x10, 0x20
SUM_COMP_FAILED, SUM_FINAL_FAILED, SUM_PHASE_CHANGED, SUM_ITER_SHIFT = 0x40, 0x80, 0x100, 16


class ParallelDesc(C.Structure):
    _fields_ = [("step", C.c_uint16), ("branches", C.c_uint16), ("allow_first", C.c_uint32)]


class Topology(C.Structure):
    _fields_ = [
        ("n_steps", C.c_uint32), ("n_edges", C.c_uint32),
        ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("step_flags", C.c_void_p),
        ("parallel", C.c_void_p), ("n_parallel", C.c_uint32),
        ("branch_allow_bits", C.c_void_p), ("n_branch_allow_bits", C.c_uint32),
    ]


class Layout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "steps_max", "words", "fields", "child_nibbles", "state_stride",
        "off_phase", "off_cond", "off_decision", "off_child", "result_stride",
        "off_ready", "off_skip", "off_fail", "off_needs_cond", "off_skip_dep", "off_phase_out")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Expansion(C.Structure):
    _fields_ = [("run", C.c_uint32), ("step", C.c_uint16), ("branch", C.c_uint16)]


class Counts(C.Structure):
    _fields_ = [("ready", C.c_uint64), ("skip", C.c_uint64), ("expansion", C.c_uint64), ("evals", C.c_uint64)]


class Batch(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_runs", C.c_uint32), ("flags", C.c_uint32), ("max_iterations", C.c_uint32),
        ("layout", Layout),
        ("state", C.c_void_p), ("result", C.c_void_p),
        ("expansion", C.c_void_p), ("expansion_cap", C.c_uint64),
        ("counts", C.c_void_p),
    ]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("arena_bytes", C.c_uint64),
                ("max_topologies", C.c_uint32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("arena_used_bytes", C.c_uint64), ("arena_cap_bytes", C.c_uint64),
                ("n_topologies", C.c_uint32), ("sm_count", C.c_uint32), ("last_grid", C.c_uint32),
                ("last_block", C.c_uint32), ("last_smem_bytes", C.c_uint32), ("last_stages", C.c_uint32),
                ("last_kernel", C.c_uint32), ("last_runs_per_trip", C.c_uint32),
                ("last_eval_chunks", C.c_uint32), ("arena_compactions", C.c_uint32)]


EVT_READY, EVT_SKIP, EVT_FAIL, EVT_NEEDS_COND, EVT_SKIP_DEP = 0x1, 0x2, 0x4, 0x8, 0x10  # BF_EVT_*


HEAD_SUMMARY_MASK, HEAD_DEAD, HEAD_LISTED, HEAD_COUNT_SHIFT = 0x7FFF, 0x7FFF, 0x8000, 16  # BF_HEAD_*
EVAL_CHANGED_ONLY = 0x10
EVAL_COUNTS_SET, EVAL_PIPELINED = 0x20, 0x40


class CompactOut(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_listed", C.c_uint32), ("head", C.c_void_p), ("events", C.c_void_p),
                ("events_cap", C.c_uint64), ("n_events", C.c_uint64)]


class SchedTables(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_stories", C.c_uint32), ("n_queues", C.c_uint32), ("global_limit", C.c_int32),
                ("global_running_base", C.c_uint32), ("story_limit", C.c_void_p), ("story_running_base", C.c_void_p),
                ("queue_limit", C.c_void_p), ("queue_aging_s", C.c_void_p), ("queue_running_base", C.c_void_p),
                ("queue_max_priority_base", C.c_void_p)]


class SchedOut(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("records", C.c_void_p), ("story_running", C.c_void_p),
                ("queue_running", C.c_void_p), ("queue_max_priority", C.c_void_p), ("global_running", C.c_void_p)]


def sched_stride(words: int) -> int:
    return (16 + 12 * words + 15) & ~15


# every symbol include/bobrafrontier.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("bf_abi_version", C.c_uint32, []),
    ("bf_strerror", C.c_char_p, [C.c_int]),
    ("bf_last_error", C.c_char_p, [C.c_void_p]),
    ("bf_create", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config)]),
    ("bf_destroy", None, [C.c_void_p]),
    ("bf_topology_put", C.c_int, [C.c_void_p, C.POINTER(Topology), C.POINTER(C.c_uint32)]),
    ("bf_topology_put_many", C.c_int, [C.c_void_p, C.POINTER(Topology), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bf_topology_put_many_checked_on_device", C.c_int, [C.c_void_p, C.POINTER(Topology), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("bf_topology_check", C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bf_topology_closure", C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bf_topology_drop", C.c_int, [C.c_void_p, C.c_uint32]),
    ("bf_topology_child_first", C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]),
    ("bf_layout_init", C.c_int, [C.POINTER(Layout), C.c_uint32, C.c_uint32, C.c_uint32]),
    ("bf_eval", C.c_int, [C.c_void_p, C.POINTER(Batch)]),
    ("bf_eval_device", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    ("bf_schedule", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(SchedTables), C.POINTER(SchedOut)]),
    ("bf_schedule_device", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(SchedTables), C.POINTER(SchedOut), C.c_void_p]),
    ("bf_resident_create", C.c_int, [C.c_void_p, C.POINTER(Layout), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bf_resident_destroy", C.c_int, [C.c_void_p, C.c_uint32]),
    ("bf_resident_upload", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    ("bf_resident_apply", C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]),
    ("bf_resident_eval", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(Counts)]),
    ("bf_resident_tick", C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(Counts)]),
    ("bf_resident_download", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    ("bf_eval_compact", C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(CompactOut)]),
    ("bf_resident_tick_compact", C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(CompactOut), C.POINTER(Counts)]),
    ("bf_group_create", C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_uint32, C.POINTER(Config)]),
    ("bf_group_destroy", None, [C.c_void_p]),
    ("bf_group_size", C.c_uint32, [C.c_void_p]),
    ("bf_group_ctx", C.c_void_p, [C.c_void_p, C.c_uint32]),
    ("bf_group_last_error", C.c_char_p, [C.c_void_p]),
    ("bf_group_shard_range", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("bf_group_topology_put_many", C.c_int, [C.c_void_p, C.POINTER(Topology), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("bf_group_eval", C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Counts)]),
    ("bf_group_schedule", C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.POINTER(SchedTables), C.POINTER(SchedOut)]),
    ("bf_alloc_pinned", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("bf_free_pinned", C.c_int, [C.c_void_p, C.c_void_p]),
    ("bf_get_stats", C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    ("bf_topology_record", C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    ("bf_topology_record_build", C.c_int, [C.POINTER(Topology), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
]

_lib = None


class FrontierError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("bobrafrontier: %s (%d): %s" % (_strerror(status), status, message))
        self.status = status


def _strerror(status: int) -> str:
    try:
        return load().bf_strerror(status).decode()
    except Exception:
        return "status"


def load() -> C.CDLL:
    """Load the CUDA extension.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "bobrapet_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the frontier path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    ab_build = bool(os.environ.get("BF_LIB"))  # an older build loaded for a same-box A/B timing: newer entry points may be absent
    for name, res, args in SYMBOLS:
        try:
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        except AttributeError:
            if ab_build:
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    if lib.bf_abi_version() != BF_ABI_VERSION and not ab_build:
        raise RuntimeError("bobrapet_b200: ABI version mismatch")
    _lib = lib
    return lib
