"""Eight reconcile workers (the reference's default MaxConcurrentReconciles, internal/config/controller_config.go:721) calling
the SAME ctx at once: bf_eval / bf_eval_compact serialise behind the ctx lock, every caller gets exactly its own batch's
records, and the serialisation costs nothing measurable against one caller doing the same work back to back (a pass over a
large batch fills the GPU by itself; for small batches the recommended shape is the single batcher of INTEGRATION.md)."""
import threading
import time

import numpy as np
import pytest

from bobrapet_b200 import _abi as A
from bobrapet_b200 import Frontier, synth
from bobrapet_b200.records import make_layout
from oracle import packed as PK

pytestmark = pytest.mark.gpu


def test_eight_concurrent_callers_get_their_own_results():
    f = Frontier(0)
    try:
        workers = 8
        S = 256
        L = make_layout(S, 0, A.F_COND | A.F_DECISION | A.F_ALL_OUT)
        jobs = []
        for w in range(workers):
            n = 2000 + 750 * w
            ts = synth.topologies(4, 100000 * w, n, S)
            slots = f.put_topologies(ts)
            state = synth.state(4, 100000 * w, n, L, slots, ts)
            want, wc = PK.evaluate(PK.PackedTopologies(ts, slots), L, state, threads=4)
            whead, wev, _ = PK.compact_events(L, want)
            jobs.append((state, want, wc, whead, wev))
        reps = 12
        errors = []

        def worker(w):
            state, want, wc, whead, wev = jobs[w]
            try:
                for i in range(reps):
                    if i % 2:
                        got, gc = f.eval(L, state)
                        assert np.array_equal(got, want) and gc == wc
                    else:
                        head, events, n_events, gc = f.eval_compact(L, state, len(wev) + 8)
                        assert n_events == len(wev) and np.array_equal(head, whead) and np.array_equal(events, wev) and gc == wc
            except Exception as e:  # noqa: BLE001
                errors.append((w, repr(e)[:300]))

        for w in range(workers):
            worker(w)                      # warm (allocations, plans)
        assert not errors, errors
        t0 = time.perf_counter()
        for w in range(workers):
            worker(w)
        serial = time.perf_counter() - t0
        threads = [threading.Thread(target=worker, args=(w,)) for w in range(workers)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        concurrent = time.perf_counter() - t0
        assert not errors, errors
        print("8 callers on one ctx: %.1f ms back to back, %.1f ms concurrently (x%.2f)" % (1e3 * serial, 1e3 * concurrent, concurrent / serial))
        assert concurrent < 1.6 * serial + 0.05
    finally:
        f.close()
