#!/usr/bin/env python
"""Regenerates tests/golden/rows_f.npz: seeded inputs and CPU-oracle outputs for the section-8(f)
rows (motion compensation, submap voxel filter, the type-1 matcher chain).  Like config1.npz these
pin the ORACLE against drift and give the GPU tests committed numbers; the reference ships no
vectors for these functions either."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import scenes  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    n = 2000
    scan = np.zeros((n, 5), np.float32)
    scan[:, :3] = rng.uniform(-60, 60, (n, 3))
    scan[:, 3] = rng.uniform(0, 255, n)
    scan[:, 4] = (np.arange(n) / (n - 1)).astype(np.float32)
    delta = np.eye(4)
    c, s = np.cos(0.09), np.sin(0.09)
    delta[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, np.cos(0.01), -np.sin(0.01)], [0, np.sin(0.01), np.cos(0.01)]])
    delta[:3, 3] = (0.8, -0.15, 0.03)
    rc, comp = O.motion_compensation(scan, delta)
    assert rc == 0
    cloud = np.zeros((6000, 5), np.float32)
    cloud[:, :3] = rng.uniform(-3, 3, (6000, 3))
    cloud[:, 3] = rng.uniform(0, 255, 6000)
    m, vox = O.voxel_grid_filter(cloud, 0.4)
    src, tgt, GT = scenes.corner_pair()
    pm = O.icp_pm_equivalent(src.astype(np.float32), tgt.astype(np.float32))
    np.savez_compressed(
        os.path.join(HERE, "rows_f.npz"), scan=scan, delta=delta, compensated=comp, cloud=cloud, voxel_size=np.float32(0.4),
        voxels=vox, pm_result=pm["result"], pm_score=pm["score"], pm_icp_fast_score=pm["icp_fast_score"],
        pm_counts=np.array([pm["n_source"], pm["n_target"], pm["iterations"], pm["score_kept"]]), pm_ok=pm["ok"])
    print("written rows_f.npz", m, pm["iterations"], pm["score"])


if __name__ == "__main__":
    main()
