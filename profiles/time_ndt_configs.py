#!/usr/bin/env python
"""Timing of the NDT (config 3) and NdtWithGicp (config 5 per-pair) alignments at BASELINE sizes:
120k-point scan -> 500k-point submap, one pair, device time via the engine's own events plus
wall clock; the CPU oracle is timed beside it.  Prints one JSON object (kept in profiles/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import staticmapping_b200 as smb  # noqa: E402


def main():
    with_oracle = "--no-oracle" not in sys.argv
    src, sub, P = bench.make_workload(0)
    s32, t32 = src.astype(np.float32), sub.astype(np.float32)
    out = {}
    for name, cls in (("ndt", smb.Ndt), ("ndt_gicp", smb.NdtWithGicp)):
        m = cls()
        m.SetInputSource(smb.InnerCloud(s32))
        m.SetInputTarget(smb.InnerCloud(t32))
        walls = []
        for _ in range(4):
            t0 = time.perf_counter()
            ok, res = m.Align(np.eye(4))
            walls.append(time.perf_counter() - t0)
        info = m.GetAlignInfo()
        E = np.linalg.inv(P) @ res
        out[name] = {"ok": ok, "wall_ms": 1e3 * float(np.median(walls[1:])), "score": m.GetFitnessScore(),
                     "info": {k: info[k] for k in ("iterations", "evaluations", "profiled_iterations", "ms_prologue",
                                                   "ms_iterations", "ms_finish", "mean_neighbors", "aux")},
                     "err_vs_truth_m": float(np.linalg.norm(E[:3, 3]))}
        if with_oracle:
            import oracle_lib as O
            t0 = time.perf_counter()
            o = O.ndt_align(s32, t32) if name == "ndt" else O.ndt_gicp_align(s32, t32)
            out[name]["oracle_s"] = time.perf_counter() - t0
            D = np.linalg.inv(o["result"]) @ res
            out[name]["parity_dt_m"] = float(np.linalg.norm(D[:3, 3]))
            out[name]["oracle_threads"] = O.num_threads()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
