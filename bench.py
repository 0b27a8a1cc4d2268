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

N_SOURCE = 120_000
N_SUBMAP = 500_000
ITERATIONS = 30
BYTES_PER_POINT_ITER = 64           # whole iteration
BYTES_KNN_PER_POINT = 40            # of which the k-NN kernel: 16 src + 16 matched + 8 write
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


def oracle_module():
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (staticmapping_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: each rank owns its pair; target prep runs on the GPU (product path) ------
    src, sub, P = make_workload(rank)
    tgt = smb.CalculateNormals(sub, device=local_rank)
    tp, tn = tgt.points, tgt.normals
    nt = tp.shape[0]
    stream = torch.cuda.Stream(device=dev)
    m = smb.IcpFast(local_rank)
    m.InitWithXml({"max_iteration": ITERATIONS, "disable_convergence_check": 1})
    m.SetStream(stream.cuda_stream)
    d_src = torch.from_numpy(src).to(dev)
    d_tp = torch.from_numpy(tp).to(dev)
    d_tn = torch.from_numpy(tn).to(dev)
    h_src = torch.from_numpy(src).pin_memory()
    h_tp = torch.from_numpy(tp).pin_memory()
    h_tn = torch.from_numpy(tn).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    guess = np.eye(4)
    launches = [0]

    def step_device():
        m.SetInputTargetDevice(d_tp.data_ptr(), d_tn.data_ptr(), nt)
        m.SetInputSourceDevice(d_src.data_ptr(), N_SOURCE)
        ok, res = m.Align(guess)
        launches[0] += m.GetAlignInfo()["kernel_launches"] + 3
        return res

    def step_host():
        m._check(m._lib.sm_set_input_target(m._h, h_tp.data_ptr(), h_tn.data_ptr(), nt), "SetInputTarget")
        m._check(m._lib.sm_set_input_source(m._h, h_src.data_ptr(), N_SOURCE), "SetInputSource")
        ok, res = m.Align(guess)
        launches[0] += m.GetAlignInfo()["kernel_launches"] + 3
        return res

    def timed(step_fn, k):
        """K steps, each bracketed by events on the engine's stream; L2 flushed in between."""
        total_ms, results = 0.0, []
        with torch.cuda.stream(stream):
            for _ in range(k):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                results.append(step_fn())
                e1.record(stream)
                e1.synchronize()
                total_ms += e0.elapsed_time(e1)
        return total_ms, results

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

    # ---- warm-up, then the timed region ----------------------------------------------------
    timed(step_device, args.warmup)
    timed(step_host, args.warmup)
    res_check = step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches[0] = 0
    ms_dev, results = timed(step_device, args.steps)
    gpu_launches = launches[0]
    # the one collective of the path: all-gather of the poses (16 doubles + score per pair)
    allgather_ms = 0.0
    if world > 1:
        poses = torch.tensor(np.stack([np.append(r.T.ravel(), 0.0) for r in results]), device=dev)
        out = [torch.empty_like(poses) for _ in range(world)]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_gather(out, poses); e1.record(); e1.synchronize()
        allgather_ms = e0.elapsed_time(e1)
    barrier()
    ms_dev_total = max_over_ranks(ms_dev + allgather_ms)
    barrier()
    ms_e2e, _ = timed(step_host, args.steps)
    barrier()
    ms_e2e_total = max_over_ranks(ms_e2e)
    # ---- dominant-kernel timing: extra profiled steps, events around every launch ----------
    m.InitWithXml({"profile_kernels": 1})
    timed(step_device, 1)
    prof = {"knn": 0.0, "accum": 0.0, "finish": 0.0, "prologue": 0.0, "n": 0}
    for _ in range(3):
        timed(step_device, 1)
        info = m.GetAlignInfo()
        prof["knn"] += info["ms_knn"]; prof["accum"] += info["ms_accum"]
        prof["finish"] += info["ms_finish"]; prof["prologue"] += info["ms_prologue"]
        prof["n"] += 1
    m.InitWithXml({"profile_kernels": 0})
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
    roofline = {"bound": "hbm", "kernel": "icp_knn_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                "avg_launch_ms": knn_ms,
                "bytes_per_launch": BYTES_KNN_PER_POINT * N_SOURCE,
                "how": "CUDA events around every launch of 3 extra profiled alignments run "
                       "right after the timed region (same stream, same inputs)",
                "per_alignment_ms": {k: prof[k] / prof["n"] for k in ("prologue", "knn", "accum", "finish")},
                "iteration_bytes_frac": (BYTES_PER_POINT_ITER * N_SOURCE * ITERATIONS /
                                         ((prof["knn"] + prof["accum"] + prof["finish"]) / prof["n"] * 1e-3)
                                         / 1e9) / peak}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(nt),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(src.nbytes + tp.nbytes + tn.nbytes + 128),
                "d2h_bytes_per_step": 128 + 400, "ms_per_step": ms_e2e_total / args.steps},
        "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline,
        "allgather_ms": allgather_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        O = oracle_module()
        ts = cpu_time_alignment(O, src, tp, tn, args.cpu_baseline_reps)
        # parity spot check of the benchmarked configuration against the oracle
        o = O.icp_fast_align(src, tp, tn, max_iteration=ITERATIONS, disable_convergence_check=True)
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
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
