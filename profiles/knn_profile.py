#!/usr/bin/env python
"""Where does the phase-A k-NN kernel spend its time?  Re-runs the search of the BASELINE
configs[1] pair with per-thread clocks (sm_debug_knn_profile) and prints per-warp / per-SM
statistics: duration distribution, rounds, the critical warp, SM busy spans."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import staticmapping_b200 as smb  # noqa: E402
from staticmapping_b200 import _lib  # noqa: E402


def main():
    src, sub, P = bench.make_workload(0)
    tgt = smb.CalculateNormals(sub)
    m = smb.IcpFast()
    m.InitWithXml({"max_iteration": 30, "disable_convergence_check": 1})
    m.SetInputTarget(tgt)
    m.SetInputSource(smb.EigenCloud(src))
    m.Align(np.eye(4))
    lib = _lib.lib()
    n = src.shape[0]
    f = lib.sm_debug_knn_profile
    f.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    out = {}
    for identity in (1, 0):
        cyc = np.zeros(n, np.uint32); rd = np.zeros(n, np.uint8); sm = np.zeros(n, np.uint8)
        t0 = np.zeros(n, np.uint64); t1 = np.zeros(n, np.uint64)
        for _ in range(3):   # warm caches; keep the last
            rc = f(m._h, identity, cyc.ctypes.data, rd.ctypes.data, sm.ctypes.data, t0.ctypes.data, t1.ctypes.data)
            assert rc == 0, rc
        nw = n // 32
        wc = cyc[: nw * 32].reshape(nw, 32).max(axis=1).astype(np.float64)      # warp duration = its slowest lane
        wr = rd[: nw * 32].reshape(nw, 32)
        base = t0.min()
        span = float(t1.max() - base) / 1e3
        per_sm_end = {int(s): float(t1[sm == s].max() - base) / 1e3 for s in np.unique(sm)}
        ends = np.array(sorted(per_sm_end.values()))
        lane_eff = float(cyc.astype(np.float64).sum() / (wc.sum() * 32))
        worst = int(np.argmax(wc))
        q = lambda a, p: float(np.percentile(a, p))
        out["first_iteration" if identity else "last_iteration"] = {
            "kernel_span_us": span,
            "warp_cycles": {"mean": float(wc.mean()), "p50": q(wc, 50), "p90": q(wc, 90), "p99": q(wc, 99), "max": float(wc.max())},
            "rounds_per_query": {"mean": float(rd.mean()), "p90": q(rd, 90), "p99": q(rd, 99), "max": int(rd.max())},
            "warp_max_rounds": {"mean": float(wr.max(axis=1).mean()), "p99": q(wr.max(axis=1), 99)},
            "cycles_per_round_worst_warp": float(wc[worst] / max(1, wr[worst].max())),
            "worst_warp": {"index": worst, "cycles": float(wc[worst]), "rounds_max": int(wr[worst].max()), "rounds_mean": float(wr[worst].mean()), "sm": int(sm[worst * 32])},
            "lane_time_efficiency": lane_eff,
            "sm_end_us": {"min": float(ends.min()), "p50": q(ends, 50), "max": float(ends.max())},
            "queries_zero_rounds_frac": float((rd == 0).mean()),
        }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
