#!/usr/bin/env python
"""Regenerates tests/golden/config1.npz: the seeded config-1 inputs (3-plane corner, 5 000 points)
and what the CPU oracle returns for them.  The reference itself ships no golden vectors for
registrators/ and cannot be built or imported here (C++/Eigen/PCL), so these fixtures pin the
ORACLE (against drift) and give the GPU tests committed numbers to compare with."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
import scenes  # noqa: E402


def main():
    src, tgt, GT = scenes.corner_pair()
    tp, tn = O.calculate_normals(tgt)
    icp = O.icp_fast_align(src, tp, tn)
    s32, t32 = src.astype(np.float32), tgt.astype(np.float32)
    ndt = O.ndt_align(s32, t32)
    ng = O.ndt_gicp_align(s32, t32)
    rng = np.random.default_rng(11)
    Q = rng.normal(size=(256, 3)) * 5.0
    ids, d2 = O.knn1(tp - tp.mean(0), Q, epsilon=3.16)
    np.savez_compressed(
        os.path.join(HERE, "config1.npz"), src=src.astype(np.float32), tgt=tgt.astype(np.float32), GT=GT,
        target_points=tp, target_normals=tn, icp_result=icp["result"], icp_score=icp["score"],
        icp_iterations=icp["iterations"], ndt_result=ndt["result"], ndt_fitness=ndt["fitness"],
        ndt_iterations=ndt["iterations"], ng_result=ng["result"], ng_score=ng["score"],
        ng_counts=np.array([ng["n_source_filtered"], ng["n_target_filtered"], ng["gicp_iterations"],
                            ng["bfgs_evaluations"]]),
        knn_query=Q, knn_ids=ids, knn_d2=d2)
    print("written", os.path.join(HERE, "config1.npz"))


if __name__ == "__main__":
    main()
