"""Frontier — thin Python handle over the C ABI (include/bobrafrontier.h).

This is the call a user of the Python binding makes; it is the same call path the
cgo shim uses (go/frontier/frontier.go): bf_create -> bf_topology_put_many ->
bf_eval / bf_eval_device.  There is no CPU path here: without the CUDA library or
without a B200 the constructor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _abi as A
from .records import EXP_DTYPE, PAR_DTYPE, TOPO_DTYPE


class TopologySet:
    """`count` topologies held as concatenated numpy arrays (keeps them alive for the C call).

    S, E, P : [count] uint32; row_ptr: concatenation of (S[i]+1) uint32 each; col_idx uint16;
    step_flags uint8; parallel PAR_DTYPE; allow_bits uint8 (bit arrays, one shared pool)."""

    def __init__(self, S, E, row_ptr, col_idx, step_flags, P=None, parallel=None, allow_bits=None):
        self.S = np.ascontiguousarray(S, dtype=np.uint32)
        self.E = np.ascontiguousarray(E, dtype=np.uint32)
        self.count = int(self.S.shape[0])
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.uint16)
        self.step_flags = np.ascontiguousarray(step_flags, dtype=np.uint8)
        self.P = np.zeros(self.count, dtype=np.uint32) if P is None else np.ascontiguousarray(P, dtype=np.uint32)
        self.parallel = np.zeros(0, dtype=PAR_DTYPE) if parallel is None else np.ascontiguousarray(parallel, dtype=PAR_DTYPE)
        self.allow_bits = np.zeros(0, dtype=np.uint8) if allow_bits is None else np.ascontiguousarray(allow_bits, dtype=np.uint8)
        assert self.row_ptr.shape[0] == int(self.S.sum()) + self.count
        assert self.col_idx.shape[0] == int(self.E.sum())
        assert self.step_flags.shape[0] == int(self.S.sum())
        assert self.parallel.shape[0] == int(self.P.sum())

    def descriptors(self) -> np.ndarray:
        """Array of bf_topology structs pointing into the concatenated arrays."""
        t = np.zeros(self.count, dtype=TOPO_DTYPE)
        S64, E64, P64 = self.S.astype(np.uint64), self.E.astype(np.uint64), self.P.astype(np.uint64)
        rp_off = np.concatenate(([0], np.cumsum(S64 + 1)[:-1])).astype(np.uint64)
        ci_off = np.concatenate(([0], np.cumsum(E64)[:-1])).astype(np.uint64)
        sf_off = np.concatenate(([0], np.cumsum(S64)[:-1])).astype(np.uint64)
        pd_off = np.concatenate(([0], np.cumsum(P64)[:-1])).astype(np.uint64)
        t["n_steps"], t["n_edges"], t["n_parallel"] = self.S, self.E, self.P
        t["row_ptr"] = np.uint64(self.row_ptr.ctypes.data) + rp_off * np.uint64(4)
        t["col_idx"] = np.uint64(self.col_idx.ctypes.data if self.col_idx.size else 0) + ci_off * np.uint64(2)
        t["step_flags"] = np.uint64(self.step_flags.ctypes.data) + sf_off
        if self.parallel.size:
            t["parallel"] = np.uint64(self.parallel.ctypes.data) + pd_off * np.uint64(PAR_DTYPE.itemsize)
        if self.allow_bits.size:
            t["branch_allow_bits"] = np.uint64(self.allow_bits.ctypes.data)
            t["n_branch_allow_bits"] = np.uint32(self.allow_bits.size * 8)
        return t

    def algorithmic_bytes(self) -> int:
        """Canonical topology bytes, SURVEY.md 8(d): 2(S+1) + 2E + S per topology."""
        return int((2 * (self.S.astype(np.int64) + 1) + 2 * self.E.astype(np.int64) + self.S.astype(np.int64)).sum())


class Frontier:
    def __init__(self, device: int = 0, arena_bytes: int = 0, _borrowed_ctx=None, reserve_sms: int = 0):
        self._lib = A.load()
        self._owned = _borrowed_ctx is None
        if _borrowed_ctx is not None:       # a shard's ctx owned by a FrontierGroup
            self._ctx = C.c_void_p(_borrowed_ctx)
            return
        self._ctx = C.c_void_p()
        cfg = A.Config(struct_size=C.sizeof(A.Config), device=device, arena_bytes=arena_bytes, max_topologies=0,
                       flags=(reserve_sms & 0xFF))   # BF_CFG_RESERVE_SMS: SMs left to other streams' kernels (the count all-gather)
        rc = self._lib.bf_create(C.byref(self._ctx), C.byref(cfg))
        if rc != A.BF_OK:
            self._ctx = None
            raise A.FrontierError(rc, "bf_create(device=%d): no CUDA fallback exists for this path" % device)

    # -- lifecycle
    def close(self):
        if getattr(self, "_ctx", None):
            if self._owned:
                self._lib.bf_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != A.BF_OK:
            raise A.FrontierError(rc, "%s: %s" % (what, self._lib.bf_last_error(self._ctx).decode()))

    # -- topologies
    def put_topologies(self, ts: TopologySet) -> np.ndarray:
        desc = ts.descriptors()
        slots = np.zeros(ts.count, dtype=np.uint32)
        rc = self._lib.bf_topology_put_many(self._ctx, desc.ctypes.data_as(C.POINTER(A.Topology)), ts.count,
                                            slots.ctypes.data_as(C.POINTER(C.c_uint32)))
        self._check(rc, "bf_topology_put_many")
        return slots

    def put_topologies_checked_on_device(self, ts: TopologySet):
        """Bulk upload with acyclicity checked by the device kernel -> (slots, status words)."""
        desc = ts.descriptors()
        slots = np.zeros(ts.count, dtype=np.uint32)
        status = np.zeros(ts.count, dtype=np.uint32)
        rc = self._lib.bf_topology_put_many_checked_on_device(
            self._ctx, desc.ctypes.data_as(C.POINTER(A.Topology)), ts.count, slots.ctypes.data_as(C.POINTER(C.c_uint32)),
            status.ctypes.data_as(C.POINTER(C.c_uint32)))
        self._check(rc, "bf_topology_put_many_checked_on_device")
        return slots, status

    def check_topologies(self, slots) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.uint32)
        status = np.zeros(slots.shape[0], dtype=np.uint32)
        self._check(self._lib.bf_topology_check(self._ctx, slots.ctypes.data_as(C.POINTER(C.c_uint32)), slots.shape[0],
                                                status.ctypes.data_as(C.POINTER(C.c_uint32))), "bf_topology_check")
        return status

    def drop_topology(self, slot: int):
        self._check(self._lib.bf_topology_drop(self._ctx, int(slot)), "bf_topology_drop")

    def child_first(self, slot: int) -> np.ndarray:
        out = np.zeros(A.MAX_PARALLEL, dtype=np.uint32)
        n = self._lib.bf_topology_child_first(self._ctx, int(slot), out.ctypes.data_as(C.POINTER(C.c_uint32)), A.MAX_PARALLEL)
        if n < 0:
            raise A.FrontierError(n, "bf_topology_child_first")
        return out[:n]

    def topology_record(self, slot: int):
        addr, nbytes = C.c_uint64(), C.c_uint32()
        self._check(self._lib.bf_topology_record(self._ctx, int(slot), C.byref(addr), C.byref(nbytes)), "bf_topology_record")
        return addr.value, nbytes.value

    # -- evaluation over HOST buffers (H2D + kernels + D2H inside the call)
    def eval(self, L: A.Layout, state: np.ndarray, result: Optional[np.ndarray] = None, flags: int = 0,
             max_iterations: int = 0, expansion_cap: int = 0):
        n = int(state.shape[0])
        assert state.dtype == np.uint8 and state.flags["C_CONTIGUOUS"] and state.shape[1] == L.state_stride
        if result is None:
            result = np.zeros((n, L.result_stride), dtype=np.uint8)
        counts = A.Counts()
        exp = np.zeros(max(expansion_cap, 1), dtype=EXP_DTYPE) if (flags & A.EVAL_EXPANSION) else None
        b = A.Batch(struct_size=C.sizeof(A.Batch), n_runs=n, flags=flags, max_iterations=max_iterations, layout=L,
                    state=state.ctypes.data, result=result.ctypes.data,
                    expansion=(exp.ctypes.data if exp is not None else None),
                    expansion_cap=(expansion_cap if exp is not None else 0),
                    counts=C.addressof(counts))
        self._check(self._lib.bf_eval(self._ctx, C.byref(b)), "bf_eval")
        cdict = {"ready": counts.ready, "skip": counts.skip, "expansion": counts.expansion, "evals": counts.evals}
        if exp is not None:
            return result, cdict, exp[:min(int(counts.expansion), expansion_cap)]
        return result, cdict




    # -- compact results: one head word per run + 16-bit (step | kind << 10) events instead of mask records
    def _compact_out(self, n_runs: int, events_cap: int, head, events):
        if head is None:
            head = np.zeros(max(n_runs, 1), dtype=np.uint32)
        if events is None:
            events = np.zeros(max(events_cap, 1), dtype=np.uint16)
        assert head.dtype == np.uint32 and head.flags["C_CONTIGUOUS"] and head.shape[0] >= n_runs
        assert events.dtype == np.uint16 and events.flags["C_CONTIGUOUS"] and events.shape[0] >= events_cap
        co = A.CompactOut(struct_size=C.sizeof(A.CompactOut), head=head.ctypes.data, events=events.ctypes.data,
                          events_cap=events_cap, n_events=0)
        return co, head, events

    def eval_compact(self, L: A.Layout, state: np.ndarray, events_cap: int, flags: int = 0, max_iterations: int = 0,
                     head: Optional[np.ndarray] = None, events: Optional[np.ndarray] = None):
        """bf_eval_compact -> (head [n] uint32, events[:min(n_events, cap)] uint16, n_events, counts)"""
        n = int(state.shape[0])
        assert state.dtype == np.uint8 and state.flags["C_CONTIGUOUS"] and state.shape[1] == L.state_stride
        co, head, events = self._compact_out(n, events_cap, head, events)
        counts = A.Counts()
        b = A.Batch(struct_size=C.sizeof(A.Batch), n_runs=n, flags=flags, max_iterations=max_iterations, layout=L,
                    state=state.ctypes.data, result=None, expansion=None, expansion_cap=0, counts=C.addressof(counts))
        self._check(self._lib.bf_eval_compact(self._ctx, C.byref(b), C.byref(co)), "bf_eval_compact")
        cdict = {"ready": counts.ready, "skip": counts.skip, "expansion": counts.expansion, "evals": counts.evals}
        return head[:n], events[:min(int(co.n_events), events_cap)], int(co.n_events), cdict

    def resident_tick_compact(self, handle: int, n_runs: int, deltas: np.ndarray, events_cap: int, flags: int = 0,
                              max_iterations: int = 0, head: Optional[np.ndarray] = None, events: Optional[np.ndarray] = None):
        """bf_resident_tick_compact -> (head, events, n_events, counts, n_listed); flags may carry A.EVAL_CHANGED_ONLY"""
        assert deltas.dtype == self.DELTA_DTYPE and deltas.flags["C_CONTIGUOUS"]
        co, head, events = self._compact_out(n_runs, events_cap, head, events)
        counts = A.Counts()
        self._check(self._lib.bf_resident_tick_compact(self._ctx, handle, deltas.ctypes.data, deltas.shape[0], n_runs, flags,
                                                       max_iterations, C.byref(co), C.byref(counts)), "bf_resident_tick_compact")
        cdict = {"ready": counts.ready, "skip": counts.skip, "expansion": counts.expansion, "evals": counts.evals}
        return head[:n_runs], events[:min(int(co.n_events), events_cap)], int(co.n_events), cdict, int(co.n_listed)

    # -- resident batches: device-side state, delta uploads (row f2)
    DELTA_DTYPE = np.dtype([("run", "<u4"), ("index", "<u2"), ("field", "u1"), ("code", "u1")])

    def resident_create(self, L: A.Layout, capacity: int) -> int:
        h = C.c_uint32()
        self._check(self._lib.bf_resident_create(self._ctx, C.byref(L), capacity, C.byref(h)), "bf_resident_create")
        return h.value

    def resident_destroy(self, handle: int):
        self._check(self._lib.bf_resident_destroy(self._ctx, handle), "bf_resident_destroy")

    def resident_upload(self, handle: int, first_run: int, state: np.ndarray):
        assert state.dtype == np.uint8 and state.flags["C_CONTIGUOUS"]
        self._check(self._lib.bf_resident_upload(self._ctx, handle, first_run, state.shape[0], state.ctypes.data), "bf_resident_upload")

    def resident_download(self, handle: int, first_run: int, n_runs: int, state_stride: int) -> np.ndarray:
        out = np.zeros((n_runs, state_stride), dtype=np.uint8)
        self._check(self._lib.bf_resident_download(self._ctx, handle, first_run, n_runs, out.ctypes.data), "bf_resident_download")
        return out

    def resident_apply(self, handle: int, deltas: np.ndarray):
        """deltas: structured array of DELTA_DTYPE (8 bytes each), at most one per (run, field, index)"""
        d = np.ascontiguousarray(deltas, dtype=self.DELTA_DTYPE)
        self._check(self._lib.bf_resident_apply(self._ctx, handle, d.ctypes.data, d.shape[0]), "bf_resident_apply")

    def resident_eval(self, handle: int, L: A.Layout, n_runs: int, result: Optional[np.ndarray] = None, flags: int = 0,
                      max_iterations: int = 0):
        if result is None:
            result = np.zeros((n_runs, L.result_stride), dtype=np.uint8)
        counts = A.Counts()
        self._check(self._lib.bf_resident_eval(self._ctx, handle, n_runs, flags, max_iterations, result.ctypes.data, C.byref(counts)),
                    "bf_resident_eval")
        return result, {"ready": counts.ready, "skip": counts.skip, "expansion": counts.expansion, "evals": counts.evals}

    def resident_tick(self, handle: int, L: A.Layout, n_runs: int, deltas: np.ndarray, result: Optional[np.ndarray] = None,
                      flags: int = 0, max_iterations: int = 0):
        """bf_resident_tick: deltas + pass + results in one call.  `deltas` must be a C-contiguous DELTA_DTYPE array
        (pinned memory from alloc_pinned makes its upload asynchronous)."""
        assert deltas.dtype == self.DELTA_DTYPE and deltas.flags["C_CONTIGUOUS"]
        if result is None:
            result = np.zeros((n_runs, L.result_stride), dtype=np.uint8)
        counts = A.Counts()
        self._check(self._lib.bf_resident_tick(self._ctx, handle, deltas.ctypes.data, deltas.shape[0], n_runs, flags, max_iterations,
                                               result.ctypes.data, C.byref(counts)), "bf_resident_tick")
        return result, {"ready": counts.ready, "skip": counts.skip, "expansion": counts.expansion, "evals": counts.evals}

    # -- redrive closure (row f3; storyrun_controller.go:535-558)
    def closure(self, slots, steps, words: int) -> np.ndarray:
        """bf_topology_closure: [count, words] uint32 masks of the steps a redrive from (slot, step) resets."""
        sl = np.ascontiguousarray(slots, dtype=np.uint32)
        st = np.ascontiguousarray(steps, dtype=np.uint32)
        assert sl.shape == st.shape
        out = np.zeros((sl.size, words), dtype=np.uint32)
        u32p = C.POINTER(C.c_uint32)
        self._check(self._lib.bf_topology_closure(self._ctx, sl.ctypes.data_as(u32p), st.ctypes.data_as(u32p), sl.size, words,
                                                  out.ctypes.data_as(u32p)), "bf_topology_closure")
        return out

    # -- limiters over the ready sets of the batch just evaluated (rows a9 / f4; dag.go:1780-1961)
    def schedule(self, L: A.Layout, n_runs: int, sched_runs: np.ndarray, story_limit, queue_limit, queue_aging_s,
                 global_limit: int = 0, story_running_base=None, queue_running_base=None, global_running_base: int = 0,
                 queue_max_priority_base=None):
        """bf_schedule: must follow eval() of the same batch.  sched_runs: [n_runs] records of 32 bytes (bf_sched_run).
        Returns dict(records [n, stride] uint8, story_running, queue_running, queue_max_priority, global_running)."""
        assert sched_runs.nbytes == 32 * n_runs and sched_runs.flags["C_CONTIGUOUS"]
        sl = np.ascontiguousarray(story_limit, dtype=np.int32)
        ql = np.ascontiguousarray(queue_limit, dtype=np.int32)
        qa = np.ascontiguousarray(queue_aging_s, dtype=np.int32)
        sb = None if story_running_base is None else np.ascontiguousarray(story_running_base, dtype=np.uint32)
        qb = None if queue_running_base is None else np.ascontiguousarray(queue_running_base, dtype=np.uint32)
        pb = None if queue_max_priority_base is None else np.ascontiguousarray(queue_max_priority_base, dtype=np.int32)
        assert ql.shape == qa.shape
        t = A.SchedTables(struct_size=C.sizeof(A.SchedTables), n_stories=sl.size, n_queues=ql.size, global_limit=global_limit,
                          global_running_base=global_running_base, story_limit=sl.ctypes.data,
                          story_running_base=(sb.ctypes.data if sb is not None else None), queue_limit=ql.ctypes.data,
                          queue_aging_s=qa.ctypes.data, queue_running_base=(qb.ctypes.data if qb is not None else None),
                          queue_max_priority_base=(pb.ctypes.data if pb is not None else None))
        stride = A.sched_stride(L.words)
        rec = np.zeros((n_runs, stride), dtype=np.uint8)
        sr = np.zeros(max(sl.size, 1), dtype=np.uint32)
        qr = np.zeros(max(ql.size, 1), dtype=np.uint32)
        mp = np.zeros(max(ql.size, 1), dtype=np.int32)
        gr = np.zeros(1, dtype=np.uint32)
        out = A.SchedOut(struct_size=C.sizeof(A.SchedOut), records=rec.ctypes.data, story_running=sr.ctypes.data,
                         queue_running=qr.ctypes.data, queue_max_priority=mp.ctypes.data, global_running=gr.ctypes.data)
        b = A.Batch(struct_size=C.sizeof(A.Batch), n_runs=n_runs, layout=L)
        self._check(self._lib.bf_schedule(self._ctx, C.byref(b), C.c_void_p(sched_runs.ctypes.data), C.byref(t), C.byref(out)),
                    "bf_schedule")
        return {"records": rec, "story_running": sr[:sl.size], "queue_running": qr[:ql.size], "queue_max_priority": mp[:ql.size],
                "global_running": int(gr[0])}

    # -- evaluation over DEVICE buffers (async on `stream`)
    def eval_device(self, L: A.Layout, n_runs: int, state_ptr: int, result_ptr: int, counts_ptr: int = 0,
                    stream: int = 0, flags: int = 0, max_iterations: int = 0, expansion_ptr: int = 0,
                    expansion_cap: int = 0):
        b = A.Batch(struct_size=C.sizeof(A.Batch), n_runs=n_runs, flags=flags, max_iterations=max_iterations, layout=L,
                    state=state_ptr, result=result_ptr, expansion=(expansion_ptr or None), expansion_cap=expansion_cap,
                    counts=(counts_ptr or None))
        self._check(self._lib.bf_eval_device(self._ctx, C.byref(b), C.c_void_p(stream)), "bf_eval_device")

    def alloc_pinned(self, nbytes: int) -> np.ndarray:
        p = C.c_void_p()
        self._check(self._lib.bf_alloc_pinned(self._ctx, nbytes, C.byref(p)), "bf_alloc_pinned")
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p.value
        return arr

    def free_pinned(self, arr: np.ndarray):
        ptr = getattr(self, "_pinned", {}).pop(arr.ctypes.data, None)
        if ptr is not None:
            self._check(self._lib.bf_free_pinned(self._ctx, C.c_void_p(ptr)), "bf_free_pinned")

    def stats(self) -> Dict[str, int]:
        s = A.Stats()
        self._check(self._lib.bf_get_stats(self._ctx, C.byref(s)), "bf_get_stats")
        return {n: getattr(s, n) for n, _ in s._fields_}


class FrontierGroup:
    """One process, several GPUs: bf_group_* (a ctx per device + an NCCL communicator over them).  Runs shard into
    contiguous blocks; the only exchange of a pass is the all-gather of the per-shard counts."""

    def __init__(self, devices: Sequence[int]):
        self._lib = A.load()
        self._g = C.c_void_p()
        dev = (C.c_int32 * len(devices))(*devices)
        rc = self._lib.bf_group_create(C.byref(self._g), dev, len(devices), None)
        if rc != A.BF_OK:
            self._g = None
            raise A.FrontierError(rc, "bf_group_create(devices=%s)" % list(devices))
        self.size = int(self._lib.bf_group_size(self._g))
        self.shards = [Frontier(_borrowed_ctx=self._lib.bf_group_ctx(self._g, k)) for k in range(self.size)]

    def close(self):
        if getattr(self, "_g", None):
            for s in self.shards:
                s.close()
            self._lib.bf_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != A.BF_OK:
            raise A.FrontierError(rc, "%s: %s" % (what, self._lib.bf_group_last_error(self._g).decode()))

    def shard_range(self, n_runs: int, shard: int):
        first, count = C.c_uint32(), C.c_uint32()
        self._check(self._lib.bf_group_shard_range(self._g, n_runs, shard, C.byref(first), C.byref(count)), "bf_group_shard_range")
        return first.value, count.value

    def put_topologies_replicated(self, ts: TopologySet) -> np.ndarray:
        desc = ts.descriptors()
        slots = np.zeros(ts.count, dtype=np.uint32)
        self._check(self._lib.bf_group_topology_put_many(self._g, desc.ctypes.data_as(C.POINTER(A.Topology)), ts.count,
                                                         slots.ctypes.data_as(C.POINTER(C.c_uint32))), "bf_group_topology_put_many")
        return slots

    def eval(self, L: A.Layout, state: np.ndarray, result: Optional[np.ndarray] = None, flags: int = 0, max_iterations: int = 0):
        """bf_group_eval -> (result, global counts, per-shard counts)"""
        n = int(state.shape[0])
        assert state.dtype == np.uint8 and state.flags["C_CONTIGUOUS"] and state.shape[1] == L.state_stride
        if result is None:
            result = np.zeros((n, L.result_stride), dtype=np.uint8)
        counts = A.Counts()
        per = (A.Counts * self.size)()
        b = A.Batch(struct_size=C.sizeof(A.Batch), n_runs=n, flags=flags, max_iterations=max_iterations, layout=L,
                    state=state.ctypes.data, result=result.ctypes.data, expansion=None, expansion_cap=0, counts=C.addressof(counts))
        self._check(self._lib.bf_group_eval(self._g, C.byref(b), per), "bf_group_eval")
        as_dict = lambda c: {"ready": c.ready, "skip": c.skip, "expansion": c.expansion, "evals": c.evals}
        return result, as_dict(counts), [as_dict(c) for c in per]

    def schedule(self, L: A.Layout, n_runs: int, sched_runs: np.ndarray, story_limit, queue_limit, queue_aging_s,
                 global_limit: int = 0, story_running_base=None, queue_running_base=None, global_running_base: int = 0,
                 queue_max_priority_base=None):
        """bf_group_schedule: must follow eval() of the same batch; limits and priority ordering hold across all shards."""
        assert sched_runs.nbytes == 32 * n_runs and sched_runs.flags["C_CONTIGUOUS"]
        sl = np.ascontiguousarray(story_limit, dtype=np.int32)
        ql = np.ascontiguousarray(queue_limit, dtype=np.int32)
        qa = np.ascontiguousarray(queue_aging_s, dtype=np.int32)
        sb = None if story_running_base is None else np.ascontiguousarray(story_running_base, dtype=np.uint32)
        qb = None if queue_running_base is None else np.ascontiguousarray(queue_running_base, dtype=np.uint32)
        pb = None if queue_max_priority_base is None else np.ascontiguousarray(queue_max_priority_base, dtype=np.int32)
        t = A.SchedTables(struct_size=C.sizeof(A.SchedTables), n_stories=sl.size, n_queues=ql.size, global_limit=global_limit,
                          global_running_base=global_running_base, story_limit=sl.ctypes.data,
                          story_running_base=(sb.ctypes.data if sb is not None else None), queue_limit=ql.ctypes.data,
                          queue_aging_s=qa.ctypes.data, queue_running_base=(qb.ctypes.data if qb is not None else None),
                          queue_max_priority_base=(pb.ctypes.data if pb is not None else None))
        rec = np.zeros((n_runs, A.sched_stride(L.words)), dtype=np.uint8)
        sr, qr = np.zeros(max(sl.size, 1), dtype=np.uint32), np.zeros(max(ql.size, 1), dtype=np.uint32)
        mp, gr = np.zeros(max(ql.size, 1), dtype=np.int32), np.zeros(1, dtype=np.uint32)
        out = A.SchedOut(struct_size=C.sizeof(A.SchedOut), records=rec.ctypes.data, story_running=sr.ctypes.data,
                         queue_running=qr.ctypes.data, queue_max_priority=mp.ctypes.data, global_running=gr.ctypes.data)
        b = A.Batch(struct_size=C.sizeof(A.Batch), n_runs=n_runs, layout=L)
        self._check(self._lib.bf_group_schedule(self._g, C.byref(b), C.c_void_p(sched_runs.ctypes.data), C.byref(t), C.byref(out)),
                    "bf_group_schedule")
        return {"records": rec, "story_running": sr[:sl.size], "queue_running": qr[:ql.size], "queue_max_priority": mp[:ql.size],
                "global_running": int(gr[0])}
