"""GPU parity of registrator::NdtWithGicp (ApproximateVoxelGrid + stock NDT + GICP/BFGS)
against the CPU oracle (oracle/gicp_oracle.cc)."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

pytestmark = pytest.mark.gpu


def _pair(pair=0):
    src, sub, P = scenes.lidar_pair(pair=pair)
    return src.astype(np.float32), sub.astype(np.float32), P


@pytest.mark.parametrize("pair", [0, 1])
def test_ndt_gicp_align_parity(pair):
    src, sub, P = _pair(pair)
    m = smb.NdtWithGicp()
    m.SetInputSource(smb.InnerCloud(src))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, res = m.Align(np.eye(4))
    o = O.ndt_gicp_align(src, sub)
    info = m.GetAlignInfo()
    assert ok and o["rc"] == 1
    # the voxel filter is an exact parallel restatement of the sequential hash-history pass
    assert info["aux"][2] == o["n_source_filtered"] and info["aux"][3] == o["n_target_filtered"]
    assert abs(info["aux"][0] - o["ndt_score"]) <= 1e-9 * max(1.0, o["ndt_score"])
    assert info["iterations"] == o["gicp_iterations"]
    assert info["profiled_iterations"] == o["bfgs_evaluations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= 1e-4 and dr <= 1e-4, (dt, dr)
    assert abs(m.GetFitnessScore() - o["score"]) < 1e-9
    gt_t, gt_r = scenes.se3_error(P, res)
    assert gt_t < 0.02 and gt_r < 2e-3      # GICP refines the coarse NDT pose


def test_without_voxel_filter_and_without_ndt():
    src, sub, P = _pair(2)
    src, sub = src[::2].copy(), sub[::3].copy()
    xml = ('<registrator_options type="3"><param name="using_voxel_filter">false</param>'
           '<param name="use_ndt">false</param></registrator_options>')
    m = smb.CreateMatcher(smb.MatcherOptions(type=smb.Type.kNdtWithGicp, registrator_options_node=xml))
    m.SetInputSource(smb.InnerCloud(src))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, res = m.Align(np.eye(4))
    o = O.ndt_gicp_align(src, sub, using_voxel_filter=False, use_ndt=False)
    assert ok and o["rc"] == 1
    assert m.GetAlignInfo()["iterations"] == o["gicp_iterations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= 1e-4 and dr <= 1e-4, (dt, dr)


def test_unknown_option_rejected():
    m = smb.NdtWithGicp()
    with pytest.raises(smb.CheckFailure):
        m.InitWithXml({"max_iteration": 3})


def test_ndt_gate_returns_false_and_guess():
    # NDT fitness > 1 -> Align returns false, result = guess, score = exp(-10) (ndt_gicp.cc:104-108)
    src, sub, _ = _pair(0)
    far = (sub + np.array([500.0, 0.0, 0.0], np.float32)).astype(np.float32)
    g = np.eye(4); g[0, 3] = 0.25
    m = smb.NdtWithGicp()
    m.SetInputSource(smb.InnerCloud(src))
    m.SetInputTarget(smb.InnerCloud(far))
    ok, res = m.Align(g)
    o = O.ndt_gicp_align(src, far, guess=g)
    assert ok is False and o["rc"] == 0
    assert np.array_equal(res, g)
    assert abs(m.GetFitnessScore() - np.exp(-10.0)) < 1e-15
    assert abs(m.GetAlignInfo()["aux"][0] - o["ndt_score"]) <= 1e-6 * o["ndt_score"]
