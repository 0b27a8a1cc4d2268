#!/usr/bin/env python
"""Section timing of icp_finish_kernel (clock64 stamps kept in the engine's state block) on the
BASELINE configs[1] pair; prints cycles between consecutive stamps of the LAST iteration and the
per-phase device times of a profiled alignment."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import staticmapping_b200 as smb  # noqa: E402
from staticmapping_b200 import _lib  # noqa: E402

NAMES = ["select_bin", "hist2_scan", "candidate_walk", "rank+tail", "warp_reduce+zero", "partials", "solve",
         "pose_update"]


def main():
    src, sub, P = bench.make_workload(0)
    tgt = smb.CalculateNormals(sub)
    m = smb.IcpFast()
    m.InitWithXml({"max_iteration": 30, "disable_convergence_check": 1, "profile_kernels": 1})
    m.SetInputTarget(tgt)
    m.SetInputSource(smb.EigenCloud(src))
    for _ in range(3):
        m.Align(np.eye(4))
    info = m.GetAlignInfo()
    lib = _lib.lib()
    f = lib.sm_debug_icp_stamps
    f.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    out = (ctypes.c_longlong * 12)()
    f(m._h, out)
    st = list(out)
    d = {NAMES[i]: st[i + 1] - st[i] for i in range(8)}
    if st[9]:
        d["solve_again_warm"] = st[9] - st[8]
    print(json.dumps({"cycles": d, "total_cycles": st[8] - st[0],
                      "ms": {k: info[k] for k in ("ms_prologue", "ms_iterations", "ms_knn", "ms_accum", "ms_finish")},
                      "kept": info["kept"]}))


if __name__ == "__main__":
    main()
