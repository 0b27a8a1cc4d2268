"""GPU parity of registrator::Ndt (pclomp NDT) against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

pytestmark = pytest.mark.gpu


def _pair(pair=0):
    src, sub, P = scenes.lidar_pair(pair=pair)
    return src.astype(np.float32), sub.astype(np.float32), P


@pytest.mark.parametrize("pair", [0, 1, 2])
def test_ndt_align_parity(pair):
    src, sub, P = _pair(pair)
    m = smb.Ndt()
    m.SetInputSource(smb.InnerCloud(src))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, res = m.Align(np.eye(4))
    o = O.ndt_align(src, sub)
    info = m.GetAlignInfo()
    assert ok and o["rc"] == 1
    assert info["iterations"] == o["iterations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= 1e-4 and dr <= 1e-4, (dt, dr)
    assert abs(m.GetFitnessScore() - o["fitness"]) <= 1e-9 * max(1.0, o["fitness"])
    # per-term math is single precision; expf (oracle, glibc) vs float(exp(double)) (device)
    # may differ by one float ulp in a handful of terms
    assert abs(info["trans_probability"] - o["trans_probability"]) < 1e-6
    assert abs(info["mean_neighbors"] - o["mean_neighbors"]) < 1e-9
    # NDT's step is clamped to [0.05, 0.1]: it is a coarse matcher
    gt_t, gt_r = scenes.se3_error(P, res)
    assert gt_t < 0.15 and gt_r < 0.05


def test_ndt_with_guess_and_strided_cloud():
    src, sub, P = _pair(3)
    guess = np.eye(4); guess[:3, 3] = [0.2, -0.1, 0.0]
    c, s = np.cos(0.02), np.sin(0.02)
    guess[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    src5 = np.zeros((src.shape[0], 5), np.float32); src5[:, :3] = src; src5[:, 3] = 7.0   # InnerPointType rows
    m = smb.Ndt()
    m.SetInputSource(smb.InnerCloud(src5))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, res = m.Align(guess)
    o = O.ndt_align(src, sub, guess=guess)
    assert m.GetAlignInfo()["iterations"] == o["iterations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= 1e-4 and dr <= 1e-4, (dt, dr)


def test_ndt_missing_cloud_returns_false():
    m = smb.Ndt()
    ok, res = m.Align(np.eye(4))          # ndt.cc:40-42
    assert ok is False


def test_ndt_rejects_any_xml_param():
    m = smb.Ndt()
    with pytest.raises(smb.CheckFailure):
        m.InitWithXml({"resolution": 2.0})


def test_create_matcher_ndt():
    m = smb.CreateMatcher(smb.MatcherOptions(type=smb.Type.kNdt))
    assert m.GetType() == smb.Type.kNdt


def test_fitness_score_with_far_and_outside_points():
    # getFitnessScore needs the EXACT nearest neighbour of every source point, however far: the grid ring
    # search must agree with the oracle's k-d tree for points several cells away from any target point
    # and for points outside the grid's bounding box
    src, sub, P = _pair(1)
    rng = np.random.default_rng(11)
    far = np.concatenate([
        sub[:50] + np.array([0.0, 0.0, 7.5], np.float32),            # a few metres above the scene
        sub[:50] + np.array([0.0, 0.0, -3.2], np.float32),           # below the ground
        rng.uniform(-1, 1, (40, 3)).astype(np.float32) * 5 + np.array([160.0, -140.0, 30.0], np.float32),   # far outside
    ])
    src2 = np.concatenate([src, far]).astype(np.float32)
    m = smb.Ndt()
    m.SetInputSource(smb.InnerCloud(src2))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, res = m.Align(np.eye(4))
    o = O.ndt_align(src2, sub)
    assert m.GetAlignInfo()["iterations"] == o["iterations"]
    assert abs(m.GetFitnessScore() - o["fitness"]) <= 1e-9 * max(1.0, o["fitness"])
    assert o["fitness"] > 10.0                                       # the far points dominate the mean: they were found
