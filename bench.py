#!/usr/bin/env python
"""bench.py — headline benchmark of the registrators/ hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one scan-pair alignment of BASELINE.json configs[1]: a synthetic 64-beam
120 000-point scan against a 500 000-point submap (106 784 target points with normals
after the caller-side CalculateNormals), point-to-plane IcpFast with the iteration count
fixed at 30 (convergence test disabled on both arms), i.e.
SetInputTarget + SetInputSource + Align including the k-d tree rebuild.

* `value`  : alignments/s with the clouds already resident in HBM (device pointers).
* `e2e`    : the same through the public host-buffer API (pinned host memory -> H2D inside
             the timed region, result read back).
* roofline : dominant kernel (the fused transform + k-NN kernel), algorithmic bytes of
             SURVEY.md section 8d over CUDA-event time measured here.
* cpu_baseline / --impl reference : the CPU oracle (a restatement of the reference's
             libnabo + Eigen path; the reference itself cannot be compiled in this image)
             on the host cores of the same box.

One process per GPU (torchrun for N > 1); every rank aligns its own pairs (weak scaling,
no data-path collective) and the resulting poses are all-gathered once over NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# more hardware work queues than the default 8, so that the per-pipeline streams do not alias
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_SOURCE = 120_000
N_SUBMAP = 500_000
ITERATIONS = 30
BYTES_PER_POINT_ITER = 64           # whole iteration
BYTES_KNN_PER_POINT = 40            # of which the k-NN kernel: 16 src + 16 matched + 8 write
# dram__bytes_read+write of icp_knn_kernel from the committed ncu --set full capture
# (profiles/r01_ncu_full_icp_knn_v2.txt, --cache-control none: warm L2, the steady state of an
# alignment; with ncu's default cache flush the same kernel reads 19.7 MB cold)
NCU_TRAFFIC_BYTES = 2_062_848
METRIC = "scan-pair alignments/sec (120k->500k pts, 30 ICP iters)"
UNIT = "alignments/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(pair: int):
    """(source (Ns,3) f64, submap (500k,3) f64, perturbation) for pair index `pair`."""
    cache = f"/tmp/sm_b200_bench_pair{pair}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        return z["src"], z["sub"], z["P"]
    from staticmapping_b200 import synth
    t0 = time.time()
    scene = synth.make_scene(0)
    sub = synth.submap(scene, seed=pair, n_points=N_SUBMAP).astype(np.float64)
    scan = synth.lidar_scan(scene, (2.0, 0.0, 0.0), seed=10_000 + pair).astype(np.float64)
    P = synth.perturbation(pair)
    src = synth.apply_se3(np.linalg.inv(P), scan + np.array([2.0, 0.0, 0.0]))
    src = src.astype(np.float32).astype(np.float64)   # clouds enter as float (InnerPointType)
    assert src.shape == (N_SOURCE, 3) and sub.shape == (N_SUBMAP, 3)
    try:
        np.savez(cache, src=src, sub=sub, P=P)
    except OSError:
        pass
    log(f"[bench] workload pair {pair} generated in {time.time() - t0:.1f}s")
    return src, sub, P


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def physical_cores():
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
        if n:
            return int(n)
    except Exception:
        pass
    return max(1, (os.cpu_count() or 2) // 2)


def oracle_module():
    """The CPU arm runs on all PHYSICAL cores: one OpenMP thread per SMT sibling was measured 3.6x
    slower on the 64-core / 128-thread box (0.51 vs 1.85 alignments/s), so unless the caller set
    OMP_NUM_THREADS the thread count is pinned before libgomp starts (both CPU legs, same rule)."""
    os.environ.setdefault("OMP_NUM_THREADS", str(physical_cores()))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib


def cpu_time_alignment(O, src, tp, tn, reps):
    """Time `reps` full oracle alignments (fixed 30 iterations, all host threads)."""
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = O.icp_fast_align(src, tp, tn, max_iteration=ITERATIONS,
                             disable_convergence_check=True)
        ts.append(time.perf_counter() - t0)
        assert r["rc"] == 1 and r["iterations"] == ITERATIONS
    return ts


def run_reference(args, rank, world):
    """--impl reference: the CPU path of the reference (oracle restatement; `oracle/_ref`
    does not exist because the reference cannot be compiled here) on the host cores."""
    if rank != 0:
        return
    O = oracle_module()
    src, sub, _ = make_workload(0)
    tp, tn = O.calculate_normals(sub)
    cpu_time_alignment(O, src, tp, tn, max(1, min(args.warmup, 1)))
    ts = cpu_time_alignment(O, src, tp, tn, args.steps)
    total = float(np.sum(ts))
    value = args.steps / total
    cores = O.num_threads()
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(tp.shape[0]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full alignments (120k->106784 pts, 30 fixed "
                                   f"iterations, k-d tree rebuilt each time), OpenMP over "
                                   f"queries with {cores} threads, rest serial as in the reference"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(n_target):
    return {"workload": "configs[1]: single 120k-pt scan -> 500k-pt submap, point-to-plane "
                        "ICP, 30 iters", "n_source": N_SOURCE, "n_submap_raw": N_SUBMAP,
            "n_target_after_prep": int(n_target), "iterations": ITERATIONS,
            "knn_epsilon": 3.16, "dist_outlier_ratio": 0.7, "convergence_check": "disabled",
            "pairs_in_flight_per_gpu": 1,
            "l2": "flushed between steps (256 MiB memset, outside the per-step events)"}


class Worker:
    """One alignment pipeline: its own matcher handle, CUDA stream and scan pair."""

    def __init__(self, smb, torch, dev, local_rank, pair):
        self.smb, self.torch = smb, torch
        self.src, sub, self.P = make_workload(pair)
        tgt = smb.CalculateNormals(sub, device=local_rank)      # target prep on the GPU
        self.tp, self.tn = tgt.points, tgt.normals
        self.nt = self.tp.shape[0]
        self.stream = torch.cuda.Stream(device=dev)
        self.m = smb.IcpFast(local_rank)
        opts = {"max_iteration": ITERATIONS, "disable_convergence_check": 1}
        if os.environ.get("SM_B200_KNN_QPC") is not None:         # A/B switch for profiles/: phase-A queries per CTA
            opts["knn_queries_per_cta"] = os.environ["SM_B200_KNN_QPC"]
        self.m.InitWithXml(opts)
        self.m.SetStream(self.stream.cuda_stream)
        self.d_src = torch.from_numpy(self.src).to(dev)
        self.d_tp = torch.from_numpy(self.tp).to(dev)
        self.d_tn = torch.from_numpy(self.tn).to(dev)
        self.h_src = torch.from_numpy(self.src).pin_memory()
        self.h_tp = torch.from_numpy(self.tp).pin_memory()
        self.h_tn = torch.from_numpy(self.tn).pin_memory()
        self.guess = np.eye(4)
        self.launches = 0
        self.last = None

    def step_device(self):
        m = self.m
        m.SetInputTargetDevice(self.d_tp.data_ptr(), self.d_tn.data_ptr(), self.nt)
        m.SetInputSourceDevice(self.d_src.data_ptr(), N_SOURCE)
        ok, res = m.Align(self.guess)
        self.launches += m.GetAlignInfo()["kernel_launches"] + 3
        self.last = (res, m.GetFitnessScore())
        return res

    def step_host(self):
        m = self.m
        m._check(m._lib.sm_set_input_target(m._h, self.h_tp.data_ptr(), self.h_tn.data_ptr(), self.nt),
                 "SetInputTarget")
        m._check(m._lib.sm_set_input_source(m._h, self.h_src.data_ptr(), N_SOURCE), "SetInputSource")
        ok, res = m.Align(self.guess)
        self.launches += m.GetAlignInfo()["kernel_launches"] + 3
        self.last = (res, m.GetFitnessScore())
        return res

    @property
    def h2d_bytes(self):
        return int(self.src.nbytes + self.tp.nbytes + self.tn.nbytes + 128)


def run_concurrent(torch, workers, steps, host, n_threads=4):
    """`steps` alignments over len(workers) pipelines.  Each of `n_threads` host threads owns a
    slice of the workers and keeps all of them in flight with sm_align_async / sm_align_wait
    (the reference gets the same concurrency from its thread pool / TBB tasks).
    Device time of the whole region: an event before (every worker stream waits on it) and
    an event after (it waits on every worker's last kernel), both on one timing stream."""
    n_threads = max(1, min(n_threads, len(workers)))
    tstream = torch.cuda.Stream()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    results = [[] for _ in workers]
    share = [len(range(w, steps, len(workers))) for w in range(len(workers))]
    errors = []
    gate = threading.Barrier(n_threads + 1)

    def collect(i):
        w = workers[i]
        ok, res = w.m.AlignWait()
        w.launches += w.m.GetAlignInfo()["kernel_launches"] + 3
        w.last = (res, w.m.GetFitnessScore())
        results[i].append((res, w.last[1]))
        w.busy = False

    def body(t):
        mine = list(range(t, len(workers), n_threads))
        left = {i: share[i] for i in mine}
        try:
            gate.wait()
            while any(left[i] > 0 for i in mine):
                for i in mine:
                    if left[i] <= 0:
                        continue
                    w = workers[i]
                    if w.busy:
                        collect(i)
                    m = w.m
                    if host:
                        m._check(m._lib.sm_set_input_target(m._h, w.h_tp.data_ptr(), w.h_tn.data_ptr(), w.nt),
                                 "SetInputTarget")
                        m._check(m._lib.sm_set_input_source(m._h, w.h_src.data_ptr(), N_SOURCE), "SetInputSource")
                    else:
                        m.SetInputTargetDevice(w.d_tp.data_ptr(), w.d_tn.data_ptr(), w.nt)
                        m.SetInputSourceDevice(w.d_src.data_ptr(), N_SOURCE)
                    m.AlignAsync(w.guess)
                    w.busy = True
                    left[i] -= 1
            for i in mine:
                if workers[i].busy:
                    collect(i)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
        for i in mine:
            workers[i].done = torch.cuda.Event()
            workers[i].done.record(workers[i].stream)

    threads = [threading.Thread(target=body, args=(t,)) for t in range(n_threads)]
    for w in workers:
        w.busy = False
    for t in threads:
        t.start()
    torch.cuda.synchronize()
    e0.record(tstream)
    for w in workers:
        w.stream.wait_event(e0)
    gate.wait()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    for w in workers:
        tstream.wait_event(w.done)
    e1.record(tstream)
    e1.synchronize()
    return e0.elapsed_time(e1), results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--inflight", type=int, default=16,
                    help="alignments in flight per GPU (the reference runs 5-6 concurrent Align "
                         "calls from its thread pool / TBB tasks)")
    ap.add_argument("--host-threads", type=int, default=8,
                    help="host threads that share the in-flight pipelines")
    ap.add_argument("--cpu-baseline-reps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import staticmapping_b200 as smb
    from staticmapping_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (staticmapping_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: every worker of every rank owns a distinct scan pair -----------------------
    P = max(1, args.inflight)
    workers = [Worker(smb, torch, dev, local_rank, rank * P + w) for w in range(P)]
    w0 = workers[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_serial(step_fn, k, stream):
        """K steps, one in flight, each bracketed by events; L2 flushed before each."""
        total_ms = 0.0
        with torch.cuda.stream(stream):
            for _ in range(k):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_fn()
                e1.record(stream)
                e1.synchronize()
                total_ms += e0.elapsed_time(e1)
        return total_ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up ---------------------------------------------------------------------------
    run_concurrent(torch, workers, args.warmup * P, False, args.host_threads)
    run_concurrent(torch, workers, args.warmup * P, True, args.host_threads)
    res_check = w0.step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    # ---- timed: device-resident inputs, P alignments in flight --------------------------------
    for w in workers:
        w.launches = 0
    ms_dev, results = run_concurrent(torch, workers, args.steps, False, args.host_threads)
    gpu_launches = sum(w.launches for w in workers)
    # the one collective of the path: all-gather of the poses (17 doubles per pair)
    allgather_ms = 0.0
    if world > 1:
        flat = [r for per in results for r in per]
        rec = parallel.pack_poses([r[0] for r in flat], [r[1] for r in flat])
        t = torch.from_numpy(rec).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_gather(out, t); e1.record(); e1.synchronize()
        allgather_ms = e0.elapsed_time(e1)
    barrier()
    ms_dev_total = max_over_ranks(ms_dev + allgather_ms)
    # ---- timed: host buffers through the public API (H2D inside) ---------------------------
    barrier()
    ms_e2e, _ = run_concurrent(torch, workers, args.steps, True, args.host_threads)
    barrier()
    ms_e2e_total = max_over_ranks(ms_e2e)
    # ---- latency: one alignment in flight, L2 flushed before each -----------------------------
    lat_steps = max(3, min(10, args.steps))
    ms_lat_dev = timed_serial(w0.step_device, lat_steps, w0.stream) / lat_steps
    ms_lat_host = timed_serial(w0.step_host, lat_steps, w0.stream) / lat_steps
    # ---- dominant-kernel timing: extra profiled steps, events around every launch ----------
    w0.m.InitWithXml({"profile_kernels": 1})
    timed_serial(w0.step_device, 1, w0.stream)
    prof = {"knn": 0.0, "accum": 0.0, "finish": 0.0, "prologue": 0.0, "n": 0}
    for _ in range(3):
        timed_serial(w0.step_device, 1, w0.stream)
        info = w0.m.GetAlignInfo()
        prof["knn"] += info["ms_knn"]; prof["accum"] += info["ms_accum"]
        prof["finish"] += info["ms_finish"]; prof["prologue"] += info["ms_prologue"]
        prof["n"] += 1
    w0.m.InitWithXml({"profile_kernels": 0})
    clocks = sampler.stop()

    value = args.steps * world / (ms_dev_total * 1e-3)
    e2e_value = args.steps * world / (ms_e2e_total * 1e-3)
    knn_ms = prof["knn"] / prof["n"] / ITERATIONS          # average launch duration
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = BYTES_KNN_PER_POINT * N_SOURCE / (knn_ms * 1e-3) / 1e9
    iter_ms = (prof["knn"] + prof["accum"] + prof["finish"]) / prof["n"]
    roofline = {"bound": "hbm", "kernel": "icp_knn_smem_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                "avg_launch_ms": knn_ms,
                "bytes_per_launch": BYTES_KNN_PER_POINT * N_SOURCE,
                "how": "CUDA events around every launch of 3 extra profiled alignments (one in "
                       "flight) run right after the timed region, same stream and inputs; the "
                       "working set is L2-resident, so this is a latency-bound kernel",
                "per_alignment_ms": {k: prof[k] / prof["n"] for k in ("prologue", "knn", "accum", "finish")},
                "iteration_bytes_frac": (BYTES_PER_POINT_ITER * N_SOURCE * ITERATIONS / (iter_ms * 1e-3) / 1e9) / peak,
                "aggregate_iteration_bytes_frac": (BYTES_PER_POINT_ITER * N_SOURCE * ITERATIONS * value / world / 1e9) / peak,
                # informational (SURVEY 8d second figure, NOT the graded fraction): bytes the traversal
                # itself touches per query — 14 nodes x 8 B + an 8-point bucket x 16 B + the 64 B above
                "traversal_inclusive_frac": (304 * N_SOURCE / (knn_ms * 1e-3) / 1e9) / peak}

    cfg = workload_config(w0.nt)
    cfg["pairs_in_flight_per_gpu"] = P
    cfg["host_threads"] = args.host_threads
    cfg["l2"] = (f"{P} distinct pairs in flight per GPU (combined working set ~{25 * P} MB vs 126 MB L2); "
                 "the latency figures flush L2 (256 MiB memset) before every step")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": w0.h2d_bytes,
                "d2h_bytes_per_step": 128 + 400, "ms_per_step": ms_e2e_total / args.steps},
        "latency": {"ms_per_alignment_device": ms_lat_dev, "ms_per_alignment_host_buffers": ms_lat_host,
                    "in_flight": 1, "icp_iterations_per_s": ITERATIONS / (iter_ms * 1e-3)},
        "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline,
        "allgather_ms": allgather_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        O = oracle_module()
        ts = cpu_time_alignment(O, w0.src, w0.tp, w0.tn, args.cpu_baseline_reps)
        # parity spot check of the benchmarked configuration against the oracle
        o = O.icp_fast_align(w0.src, w0.tp, w0.tn, max_iteration=ITERATIONS, disable_convergence_check=True)
        E = np.linalg.inv(o["result"]) @ res_check
        out["parity_vs_oracle"] = {"dt_m": float(np.linalg.norm(E[:3, 3])),
                                   "dr_rad": float(np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)))}
        cores = O.num_threads()
        out["cpu_baseline"] = {
            "value": len(ts) / float(np.sum(ts)), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(ts)} full alignments of the same workload (30 fixed iterations, tree "
                      f"rebuilt each time), OpenMP over queries with {cores} threads"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    # orderly teardown: drain the GPU and destroy the engine handles before the interpreter
    # (and torch's CUDA context) goes away
    torch.cuda.synchronize()
    for w in workers:
        w.m.SetStream(0)
        w.m.__del__()
    del workers, w0
    torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
