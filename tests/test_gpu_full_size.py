"""GPU parity at the FULL sizes of BASELINE.json (120 000-point scan -> 500 000-point submap) for the
three matchers, against the CPU oracle: poses within 1e-4 m / 1e-4 rad (north_star), equal iteration
counts.  The oracle needs ~0.5 s (IcpFast), ~0.4 s (Ndt) and ~1.2 s (NdtWithGicp) per alignment."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

pytestmark = pytest.mark.gpu

TOL_T = 1e-4   # metres
TOL_R = 1e-4   # radians


@pytest.fixture(scope="module")
def pair0():
    return scenes.full_size_pair(0)


@pytest.mark.parametrize("fixed", [False, True])
def test_config2_icp_full_size(pair0, fixed):
    src, sub, P = pair0
    tp, tn = O.calculate_normals(sub)
    assert tp.shape[0] == 106_784                        # 500 000 -> 106 784 leaves
    # target prep on the GPU gives the same decimated cloud
    g = smb.CalculateNormals(sub)
    assert g.points.shape[0] == tp.shape[0]
    m = smb.IcpFast()
    m.InitWithXml({"max_iteration": 30, "disable_convergence_check": int(fixed)})
    m.SetInputSource(smb.EigenCloud(src)); m.SetInputTarget(smb.EigenCloud(tp, tn))
    ok, res = m.Align(np.eye(4))
    o = O.icp_fast_align(src, tp, tn, max_iteration=30, disable_convergence_check=fixed)
    info = m.GetAlignInfo()
    assert ok and o["rc"] == 1
    assert info["iterations"] == o["iterations"]
    if fixed:
        assert info["iterations"] == 30
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)
    assert abs(m.GetFitnessScore() - o["score"]) < 1e-9
    assert info["kept"] == 84_000                        # int(120 000 * (double)0.7f) + 1 matches at or below the limit
    gt_t, gt_r = scenes.se3_error(P, res)
    # the alignment is real (the convergence test stops at 1e-2 m mean step: not noise-limited)
    assert gt_t < (5e-3 if fixed else 3e-2) and gt_r < (1e-3 if fixed else 5e-3), (gt_t, gt_r)


def test_config2_knn_on_real_clouds_bit_exact(pair0):
    # the raw k-NN on the benchmark's own clouds (lidar geometry, not Gaussian blobs), eps = 3.16 and 0
    src, sub, P = pair0
    tp, tn = O.calculate_normals(sub)
    for eps in (3.16, 0.0):
        ids_o, d2_o = O.knn1(tp, src, epsilon=eps)
        ids_g, d2_g = smb.knn1(tp, src, epsilon=eps)
        assert np.array_equal(ids_g, ids_o)
        assert np.array_equal(d2_g, d2_o)


def test_config3_ndt_full_size(pair0):
    src, sub, P = pair0
    s32, t32 = src.astype(np.float32), sub.astype(np.float32)
    m = smb.Ndt()
    m.SetInputSource(smb.InnerCloud(s32)); m.SetInputTarget(smb.InnerCloud(t32))
    ok, res = m.Align(np.eye(4))
    o = O.ndt_align(s32, t32)
    info = m.GetAlignInfo()
    assert ok and o["rc"] == 1
    assert info["iterations"] == o["iterations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)
    assert abs(m.GetFitnessScore() - o["fitness"]) <= 1e-9 * max(1.0, o["fitness"])
    assert abs(info["mean_neighbors"] - o["mean_neighbors"]) < 1e-9
    assert abs(info["trans_probability"] - o["trans_probability"]) < 1e-6


def test_config5_ndt_gicp_pair_full_size(pair0):
    src, sub, P = pair0
    s32, t32 = src.astype(np.float32), sub.astype(np.float32)
    m = smb.NdtWithGicp()
    m.SetInputSource(smb.InnerCloud(s32)); m.SetInputTarget(smb.InnerCloud(t32))
    ok, res = m.Align(np.eye(4))
    o = O.ndt_gicp_align(s32, t32)
    info = m.GetAlignInfo()
    assert ok and o["rc"] == 1
    assert info["aux"][2] == o["n_source_filtered"] and info["aux"][3] == o["n_target_filtered"]
    assert info["iterations"] == o["gicp_iterations"]
    assert info["profiled_iterations"] == o["bfgs_evaluations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)
    assert abs(m.GetFitnessScore() - o["score"]) < 1e-9
