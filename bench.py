#!/usr/bin/env python
"""bench.py — step ready-evaluations/sec of the StoryRun DAG frontier pass.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 3]

A "step" is one frontier pass over one batch of synthetic StoryRuns: BASELINE.json
configs[2] — 100k StoryRuns x 256 steps, random DAG with in-degree 4 — per GPU (weak
scaling: every rank evaluates its own 100k-run shard; the only cross-GPU traffic is one
NCCL all-gather of the per-shard counts per pass).  One evaluation = one (StoryRun, step)
visit of the findReadySteps loop (dag.go:2647).

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in
HBM; `e2e` = the same metric through the public host-buffer call bf_eval (pinned host
buffers, H2D + kernel + D2H inside the timed region); `roofline` relates the kernel's
algorithmic bytes to the measured HBM copy peak; `cpu_baseline` times the CPU oracle on the
box's host cores (a reported baseline, not the target).

--impl reference times the reference's algorithm on the CPU.  The reference is Go and there
is no Go toolchain in this image (nor network), so the CPU arm is the oracle restatement
(kind "port"), with all host threads, on bounded samples of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "step ready-evals/sec @100k StoryRuns x 256 steps"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config index (2..5), default 3 = configs[2]")
    ap.add_argument("--runs", type=int, default=0, help="StoryRuns per GPU (default: the config's N, capped for cfg 5)")
    ap.add_argument("--rot", type=int, default=3, help="disjoint input copies rotated between passes (L2 hygiene)")
    ap.add_argument("--cpu-sample-runs", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shared", type=int, default=0, help="shared-topology mode: D distinct topologies (0 = unique)")
    ap.add_argument("--ncu", action="store_true", help="profiling run: few passes, no e2e/cpu legs")
    ap.add_argument("--no-graph", action="store_true", help="launch every pass eagerly instead of replaying a CUDA graph")
    ap.add_argument("--unroll", type=int, default=0, help="passes captured per CUDA graph (default 4*rot)")
    return ap.parse_args()


CFG_N = {2: 10_000, 3: 100_000, 4: 100_000, 5: 125_000}
CFG_S = {2: 64, 3: 256, 4: 256, 5: 1024}


def algorithmic_bytes(cfg, n_runs, S, E, n_topo, child_nibbles, n_expansion):
    """N*S*(0.75+m) + D*(2(S+1)+2E+S) + N*P*B*0.5 + 8*X."""
    m = 0.5 if cfg in (4,) else 0.0
    state = n_runs * S * (0.75 + m)
    topo = n_topo * (2 * (S + 1) + 2 * E + S)
    child = n_runs * child_nibbles * 0.5
    return state + topo + child + 8 * n_expansion


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        self.max_mhz = 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        for bit, n in names.items():
            if r & bit:
                self.reasons.add(n)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag:
            try:
                self.sample()
            except Exception:
                break
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz or None, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_arm(args, cfg, S, sample_runs, threads, reps, impl="refshape"):
    """Time a CPU restatement of the reference's per-iteration work on a bounded sample of the workload.

    impl="refshape": oracle/refshape.cc — the reference's own data shapes (string-keyed maps, dependency graph
    rebuilt per pass, buildStateMaps as often as dag.go calls it); impl="packed": oracle/packed_ref.c (bitmask)."""
    from bobrapet_b200 import _abi as A, synth
    from bobrapet_b200.records import make_layout
    from oracle import packed as PK
    ts = synth.topologies(cfg, 0, sample_runs, S)
    pt = PK.PackedTopologies(ts)
    child = pt.max_child_nibbles()
    fields = (A.F_COND | A.F_DECISION if cfg in (4, 5) else 0) | (A.F_CHILD if child else 0)
    L = make_layout(S, child, fields)
    st = synth.state(cfg, 0, sample_runs, L, np.arange(sample_runs, dtype=np.uint32), ts,
                     pt.child_first[:int(ts.P[0])] if child else None)
    times = []
    if impl == "refshape":
        rs = PK.RefShapeBatch(pt, L, st)  # object construction is untimed (the informer cache holds objects)
        rs.run(threads)  # warm
        evals = 0
        for _ in range(reps):
            t0 = time.perf_counter()
            _, evals = rs.run(threads)
            times.append(time.perf_counter() - t0)
        rs.close()
        return evals, times
    PK.evaluate(pt, L, st, 0, 0, threads)  # warm
    for _ in range(reps):
        t0 = time.perf_counter()
        _, counts = PK.evaluate(pt, L, st, 0, 0, threads)
        times.append(time.perf_counter() - t0)
    return counts["evals"], times


def main():
    args = parse()
    cfg = args.config
    S = CFG_S[cfg]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = world if world > 1 else 1
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample_runs or min(20_000, 250 * cores)
        for _ in range(max(args.warmup, 0)):
            pass  # warm-up happens inside cpu_arm (one untimed pass)
        evals, times = cpu_arm(args, cfg, S, sample, cores, max(args.steps, 1))
        dt = float(np.sum(times))
        v = evals * len(times) / dt
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
                "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / len(times), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 bitmask", "data": "synthetic",
                "config": {"workload": "cfg%d: %d StoryRuns x %d steps sample of BASELINE configs[%d]" % (cfg, sample, S, cfg - 1),
                           "impl_note": "reference is Go (no toolchain here): reference-shaped C++ restatement oracle/refshape.cc"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "oracle/refshape.cc, %d StoryRuns x %d steps per step, %d threads" % (sample, S, cores)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from bobrapet_b200 import _abi as A, Frontier, synth
    from bobrapet_b200.records import make_layout
    from bobrapet_b200.sharding import CountExchange, global_offsets

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_runs = args.runs or CFG_N[cfg]
    if args.ncu:
        args.steps, args.warmup, args.no_e2e, args.no_cpu = min(args.steps, 3), min(args.warmup, 3), True, True
    ROT = max(1, args.rot)

    fr = Frontier(local_rank)
    run_lo = rank * n_runs
    fields = A.F_COND | A.F_DECISION if cfg in (4, 5) else 0
    sets = []
    n_topo = args.shared or n_runs
    child = 0
    E = None
    for k in range(ROT):
        ts = synth.topologies(cfg, run_lo + k * 10_000_019, n_topo, S)
        E = int(ts.E[0])
        slots = fr.put_topologies(ts)
        cf = fr.child_first(int(slots[0])) if int(ts.P[0]) else None
        if cf is not None:
            child = int((int(cf[-1]) + int(ts.parallel["branches"][int(ts.P[0]) - 1]) + 7) // 8 * 8)
        L = make_layout(S, child, fields | (A.F_CHILD if child else 0))
        run_slots = slots if not args.shared else slots[np.arange(n_runs) % n_topo]
        ts_state = ts if not args.shared else synth.topologies(cfg, run_lo + k * 10_000_019, n_runs, S)
        st = synth.state(cfg, run_lo + k * 10_000_019, n_runs, L, run_slots, ts_state, cf)
        d_state = torch.from_numpy(st).to(dev)
        d_result = torch.zeros((n_runs, L.result_stride), dtype=torch.uint8, device=dev)
        d_counts = torch.zeros(4, dtype=torch.int64, device=dev)
        sets.append((L, d_state, d_result, d_counts, st))
        del ts
    L = sets[0][0]
    stream = torch.cuda.current_stream()
    exch = CountExchange(dev, world)
    gathered = [exch.new_buffer() for _ in range(ROT)]
    launches = [0]

    def one_pass(i, st_):
        Lk, d_state, d_result, d_counts, _ = sets[i % ROT]
        d_counts.zero_()
        fr.eval_device(Lk, n_runs, d_state.data_ptr(), d_result.data_ptr(), d_counts.data_ptr(), st_.cuda_stream)
        launches[0] += 1
        if world > 1:
            # the path's one collective: all-gather of the per-shard counts, overlapped with the next pass
            exch.gather(d_counts, gathered[i % ROT], st_)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    work_stream = torch.cuda.Stream()
    use_graph = not (args.no_graph or args.ncu)
    U = args.unroll or 4 * ROT
    U = max(ROT, (U // ROT) * ROT)
    graph = None
    with torch.cuda.stream(work_stream):
        for i in range(max(args.warmup, ROT)):
            one_pass(i, work_stream)
        exch.join(work_stream)
    barrier()
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=work_stream, capture_error_mode="thread_local"):
                for i in range(U):
                    one_pass(i, work_stream)
                exch.join(work_stream)
            with torch.cuda.stream(work_stream):
                graph.replay()  # one untimed replay
            barrier()
        except Exception as e:  # capture unsupported: fall back to eager launches
            sys.stderr.write("bench: CUDA graph capture failed (%s); eager launches\n" % e)
            graph = None
            barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches[0] = 0
    with torch.cuda.stream(work_stream):
        e0.record(work_stream)
        done = 0
        if graph is not None:
            while done + U <= args.steps:
                graph.replay()
                done += U
                launches[0] += U
        for i in range(done, args.steps):
            one_pass(i, work_stream)
        exch.join(work_stream)
        e1.record(work_stream)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    stream = torch.cuda.current_stream()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    evals_per_pass = n_runs * S * n_gpus
    value = evals_per_pass * args.steps / (ms * 1e-3)
    counts_host = sets[(args.steps - 1) % ROT][3].cpu().numpy().tolist()
    offsets = global_offsets(gathered[(args.steps - 1) % ROT], rank) if world > 1 else None
    timed_launches = launches[0]

    # ---- kernel-only duration for the roofline: event pair around each launch
    kdur = []
    for i in range(min(args.steps, 40)):
        Lk, d_state, d_result, d_counts, _ = sets[i % ROT]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fr.eval_device(Lk, n_runs, d_state.data_ptr(), d_result.data_ptr(), d_counts.data_ptr(), stream.cuda_stream,
                       flags=A.EVAL_NO_COUNTS)
        b.record(stream)
        torch.cuda.synchronize()
        kdur.append(a.elapsed_time(b))
    k_ms = float(np.median(kdur))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    abytes = algorithmic_bytes(cfg, n_runs, S, E, n_topo, child, counts_host[2] if cfg == 5 else 0)
    # The dominant kernel's average launch duration over the timed region: the region holds, per pass, one
    # frontier_kernel launch and one 32-byte counter fill (plus the overlapped count all-gather at N > 1), so
    # region time / launches is an UPPER bound on the kernel's duration and the fraction below a lower bound.  The
    # event-bracketed single launches above (kernel_ms_isolated) carry per-launch gaps and are reported beside it.
    region_ms = ms / max(timed_launches, 1)
    kernel_ms = min(region_ms, k_ms)
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get("cfg%d" % cfg, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    st_stats = fr.stats()

    # ---- e2e: the public host-buffer call, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        Lk = sets[0][0]
        hs = fr.alloc_pinned(n_runs * Lk.state_stride).reshape(n_runs, Lk.state_stride)
        hr = fr.alloc_pinned(n_runs * Lk.result_stride).reshape(n_runs, Lk.result_stride)
        hs[:] = sets[0][4]
        for _ in range(3):
            fr.eval(Lk, hs, hr)
        barrier()
        k_e2e = max(5, min(args.steps, 50))
        t0 = time.perf_counter()
        for i in range(k_e2e):
            _, c = fr.eval(Lk, hs, hr)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": evals_per_pass * k_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": int(n_runs * Lk.state_stride),
               "d2h_bytes_per_step": int(n_runs * Lk.result_stride + 32), "steps": k_e2e,
               "api": "bf_eval (host buffers, synchronous): H2D state + frontier kernel + D2H results/counts per step"}
        # informational: the incremental path of row f2 — state resident on the device, a tick sends only deltas
        # (here 1 % of all (run, step) phase codes change per tick) and reads every result record back
        try:
            rng = np.random.default_rng(1234 + rank)
            hres = fr.resident_create(Lk, n_runs)
            fr.resident_upload(hres, 0, sets[0][4])
            k_delta = max(1, (n_runs * S) // 100)
            dsets = []
            for _ in range(3):
                flat = rng.choice(n_runs * S, size=k_delta, replace=False)
                d = fr.alloc_pinned(k_delta * 8).view(fr.DELTA_DTYPE)       # pinned: the delta upload is asynchronous
                d["run"], d["index"], d["field"] = flat // S, flat % S, A.DELTA_PHASE
                d["code"] = rng.choice([0, 2, 3, 3, 3, 4, 13], size=k_delta)
                dsets.append(d)
            for i in range(3):
                fr.resident_tick(hres, Lk, n_runs, dsets[i], hr)
            barrier()
            t0 = time.perf_counter()
            for i in range(k_e2e):
                fr.resident_tick(hres, Lk, n_runs, dsets[i % 3], hr)
            dti = time.perf_counter() - t0
            tti = torch.tensor([dti], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tti, op=dist.ReduceOp.MAX)
            e2e["incremental"] = {"value": evals_per_pass * k_e2e / float(tti.item()), "unit": UNIT, "change_rate": 0.01,
                                  "h2d_bytes_per_step": int(k_delta * 8), "d2h_bytes_per_step": int(n_runs * Lk.result_stride + 32),
                                  "api": "bf_resident_tick (deltas + pass + results, one call): state stays on the device (row f2)"}
            fr.resident_destroy(hres)
            for d in dsets:
                fr.free_pinned(d.view(np.uint8))
        except Exception as ex:  # never lose the contract line over the informational leg
            e2e["incremental"] = {"error": str(ex)[:200]}
        fr.free_pinned(hs.reshape(-1))
        fr.free_pinned(hr.reshape(-1))

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        sample = args.cpu_sample_runs or min(20_000, 250 * cores)
        evals, times = cpu_arm(args, cfg, S, sample, cores, 5, "refshape")
        ev8, t8 = cpu_arm(args, cfg, S, sample, min(8, cores), 5, "refshape")
        evp, tp = cpu_arm(args, cfg, S, sample, cores, 5, "packed")
        cpu = {"value": evals / float(np.median(times)), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle/refshape.cc (reference-shaped: string-keyed maps, per-pass graph rebuild) on %d StoryRuns x %d steps, "
                         "%d threads, median of 5; the Go reference itself cannot be built here" % (sample, S, cores),
               "at_8_threads": ev8 / float(np.median(t8)),   # the reference's default MaxConcurrentReconciles
               "packed_cpu": {"value": evp / float(np.median(tp)), "threads": cores, "impl": "oracle/packed_ref.c (bitmask)"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32 bitmask (integer)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[%d]: %d StoryRuns x %d steps per GPU, %s, E=%d/topology" %
                       (cfg - 1, n_runs, S, ("unique topology per run" if not args.shared else "%d shared topologies" % n_topo), E),
                       "l2": "inputs %.0f MB/pass > 126 MB L2, rotated over %d disjoint copies" % (abytes / 1e6, ROT),
                       "parallelism": "runs sharded across %d GPU(s); one NCCL all-gather of counts per pass" % n_gpus,
                       "grid": st_stats["last_grid"], "block": st_stats["last_block"], "smem": st_stats["last_smem_bytes"],
                       "stages": st_stats["last_stages"], "launch": ("cuda-graph x%d passes" % U) if graph is not None else "eager"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel_ms": kernel_ms, "kernel_ms_isolated": k_ms, "algorithmic_bytes_per_launch": abytes,
                         "peak_source": peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": timed_launches, "clocks": sampler.result(),
            "global_counts_last_pass": (offsets["total"] if offsets else None),
            "counts_last_pass": {"ready": counts_host[0], "skip": counts_host[1], "expansion": counts_host[2], "evals": counts_host[3]},
        }
        print(json.dumps(line), flush=True)
    # tear-down: release the captured graph (it holds NCCL work) before the process group; a rank that
    # lingers here would only burn GPU time, so leave hard once everything is flushed.
    graph = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    fr.close()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        try:
            dist.barrier(device_ids=[local_rank])
        except Exception:
            pass
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
