"""GPU parity of the IcpUsingPointMatcher stand-in (type 1, icp_pointmatcher.cc:125-247) against
the same chain assembled from the CPU oracle's pieces (tests/oracle_lib.icp_pm_equivalent):
CalculateNormals on the reference cloud, hash subsample of the reading cloud, IcpFast::Align with
the 150-iteration counter, Align() false below score 0.6."""
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
def test_matches_oracle_chain(pair):
    src, sub, P = _pair(pair)
    m = smb.IcpUsingPointMatcher()
    m.SetInputSource(smb.InnerCloud(src))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, res = m.Align(np.eye(4))
    o = O.icp_pm_equivalent(src, sub)
    info = m.GetAlignInfo()
    assert o["rc"] == 1 and ok == o["ok"]
    assert int(info["aux"][2]) == o["n_source"] and int(info["aux"][3]) == o["n_target"]   # identical filters
    assert info["iterations"] == o["iterations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= 1e-4 and dr <= 1e-4, (dt, dr)            # the parity bar; observed ~1e-13
    assert abs(m.GetFitnessScore() - o["score"]) < 1e-9              # score of the unfiltered clouds after the result
    assert abs(info["aux"][0] - o["icp_fast_score"]) < 1e-9 and int(info["aux"][1]) == o["score_kept"]
    assert m.GetFitnessScore() > o["icp_fast_score"]                  # the raw reference is 5x denser than the filtered one
    gt_t, gt_r = scenes.se3_error(P, res)
    assert gt_t < 0.05 and gt_r < 0.01                      # and it actually registers the scan


def test_create_matcher_type_1_and_repeat_align():
    src, sub, _ = _pair(2)
    m = smb.CreateMatcher(smb.MatcherOptions(type=smb.Type.kIcpPM))
    assert isinstance(m, smb.IcpUsingPointMatcher) and m.GetType() == smb.Type.kIcpPM
    m.SetInputSource(smb.InnerCloud(src))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok1, r1 = m.Align(np.eye(4))
    ok2, r2 = m.Align(np.eye(4))                            # filtered clouds are cached
    assert ok1 == ok2 and np.array_equal(r1, r2)
    m.SetInputSource(smb.InnerCloud(src[: src.shape[0] // 2]))   # new reading cloud only
    ok3, r3 = m.Align(np.eye(4))
    o = O.icp_pm_equivalent(src[: src.shape[0] // 2], sub)
    dt, dr = scenes.se3_error(o["result"], r3)
    assert dt <= 1e-4 and dr <= 1e-4 and ok3 == o["ok"]


def test_low_score_returns_false_and_no_xml_option():
    src, sub, _ = _pair(0)
    far = src.copy()
    far[:, 0] += 15.0                                        # a hopeless guess: score < 0.6
    m = smb.IcpUsingPointMatcher()
    m.SetInputSource(smb.InnerCloud(far))
    m.SetInputTarget(smb.InnerCloud(sub))
    ok, _ = m.Align(np.eye(4))
    o = O.icp_pm_equivalent(far, sub)
    assert ok == o["ok"] and abs(m.GetFitnessScore() - o["score"]) < 1e-9
    assert (m.GetFitnessScore() >= 0.6) == ok
    with pytest.raises(smb.CheckFailure):
        m.InitWithXml({"max_iteration": 10})                # registers no option (interface.cc:66)
    m.SetEngineOptions(reading_sample_prob=1.0)              # keep every reading point
    m.SetInputSource(smb.InnerCloud(src))
    ok, res = m.Align(np.eye(4))
    o = O.icp_pm_equivalent(src, sub, prob=1.0)
    assert int(m.GetAlignInfo()["aux"][2]) == src.shape[0]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= 1e-4 and dr <= 1e-4
