#!/usr/bin/env python
"""Wall time per Align of Ndt / NdtWithGicp at full BASELINE sizes (one in flight), with the engine's own
event timings (grid build / iterations / fitness)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import staticmapping_b200 as smb
src, sub, _ = bench.make_workload(0)
s32, t32 = src.astype(np.float32), sub.astype(np.float32)
WARM = int(sys.argv[1]) if len(sys.argv) > 1 else 3
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
out = {}
for name, cls in (("ndt", smb.Ndt), ("ndt_gicp", smb.NdtWithGicp)):
    m = cls(0)
    m.SetInputSource(smb.InnerCloud(s32)); m.SetInputTarget(smb.InnerCloud(t32))
    for _ in range(WARM):
        m.Align(np.eye(4))
    ts = []
    for _ in range(REPS):
        t0 = time.perf_counter(); m.Align(np.eye(4)); ts.append(time.perf_counter() - t0)
    info = m.GetAlignInfo()
    out[name] = {"ms_per_align_median": 1e3 * float(np.median(ts)), "ms_min": 1e3 * min(ts),
                 "ms_prologue": info["ms_prologue"], "ms_iterations": info["ms_iterations"], "ms_finish": info["ms_finish"],
                 "iterations": info["iterations"], "evaluations": info["evaluations"], "aux": info["aux"]}
    print(name, out[name], flush=True)
print(json.dumps(out))
