"""Known-answer checks of the motion-compensation restatement (oracle/motion_oracle.cc;
reference: builder/map_builder.cc:232-257, common/math.h:198-211, common/math.cc:177-195) and of
the host-side AverageTransforms mirror.  No GPU."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation, Slerp

import oracle_lib as O
from staticmapping_b200 import registrators as R


def se3(rpy_deg, t):
    T = np.eye(4)
    T[:3, :3] = Rotation.from_euler("xyz", rpy_deg, degrees=True).as_matrix()
    T[:3, 3] = t
    return T


DELTA = se3((1.5, -2.0, 8.0), (0.9, -0.2, 0.05))


def test_interpolate_endpoints_and_midpoint():
    rc, T0 = O.interpolate_transform(np.eye(4), DELTA, 0.0)
    assert rc == 0 and np.allclose(T0, np.eye(4), atol=1e-15)
    rc, T1 = O.interpolate_transform(np.eye(4), DELTA, 1.0)
    assert np.allclose(T1, DELTA, atol=1e-12)
    rc, Th = O.interpolate_transform(np.eye(4), DELTA, 0.5)
    # slerp midpoint == rotation by half the angle about the same axis; translation halves
    rv = Rotation.from_matrix(DELTA[:3, :3]).as_rotvec()
    assert np.allclose(Th[:3, :3], Rotation.from_rotvec(0.5 * rv).as_matrix(), atol=1e-12)
    assert np.allclose(Th[:3, 3], 0.5 * DELTA[:3, 3], atol=1e-15)


def test_interpolate_between_two_poses_matches_scipy_slerp():
    A = se3((10, 5, -30), (1, 2, 3))
    B = se3((-20, 15, 40), (-1, 0, 2))
    s = Slerp([0, 1], Rotation.from_matrix(np.stack([A[:3, :3], B[:3, :3]])))
    for f in (0.125, 0.3, 0.77):
        rc, T = O.interpolate_transform(A, B, f)
        assert rc == 0
        assert np.allclose(T[:3, :3], s([f])[0].as_matrix(), atol=1e-12)
        assert np.allclose(T[:3, 3], A[:3, 3] + (B[:3, 3] - A[:3, 3]) * np.float32(f), atol=1e-12)


def test_interpolate_check_failure_outside_unit_interval():
    assert O.interpolate_transform(np.eye(4), DELTA, 1.5)[0] == -1
    assert O.interpolate_transform(np.eye(4), DELTA, -0.1)[0] == -1


def test_small_rotation_takes_the_lerp_branch():
    # |dot| >= 1 - eps: Eigen's slerp degrades to linear weights (identity rotation here)
    D = np.eye(4)
    D[:3, 3] = (0.3, 0.0, 0.0)
    rc, T = O.interpolate_transform(np.eye(4), D, 0.25)
    assert np.allclose(T[:3, :3], np.eye(3), atol=0) and np.allclose(T[:3, 3], (0.075, 0, 0), atol=1e-9)


def test_motion_compensation_points():
    rng = np.random.default_rng(3)
    n = 2000
    pts = np.zeros((n, 5), np.float32)
    pts[:, :3] = rng.uniform(-50, 50, (n, 3))
    pts[:, 3] = rng.uniform(0, 255, n)
    pts[:, 4] = np.linspace(0, 1, n, dtype=np.float32)
    rc, out = O.motion_compensation(pts, DELTA)
    assert rc == 0
    assert np.array_equal(out[:, 3:], pts[:, 3:])                    # intensity, factor copied
    assert np.array_equal(out[0, :3], pts[0, :3])                    # factor 0: untouched
    full = (DELTA[:3, :3] @ pts[-1, :3].astype(np.float64) + DELTA[:3, 3]).astype(np.float32)
    assert np.allclose(out[-1, :3], full, rtol=0, atol=1e-5)         # factor 1: the whole delta
    rv = Rotation.from_matrix(DELTA[:3, :3]).as_rotvec()
    for i in (137, 999, 1500):
        f = np.float64(pts[i, 4])
        want = Rotation.from_rotvec(f * rv).as_matrix() @ pts[i, :3].astype(np.float64) + f * DELTA[:3, 3]
        assert np.allclose(out[i, :3], want, atol=1e-5)
    bad = pts.copy()
    bad[5, 4] = 1.25
    assert O.motion_compensation(bad, DELTA)[0] == -1


def test_average_transforms_oracle_and_mirror():
    A = se3((1, -2, 10), (1.0, 0.5, 0.1))
    B = se3((3, 2, 14), (1.4, 0.1, 0.3))
    rc, M = O.average_transforms([A, B])
    assert rc == 0
    want = se3((2, 0, 12), (1.2, 0.3, 0.2))
    assert np.allclose(M, want, atol=1e-12)
    assert np.allclose(R.AverageTransforms([A, B]), M, atol=1e-14)
    assert np.allclose(R.AverageTransforms([A]), A, atol=1e-14)
    with pytest.raises(R.CheckFailure):
        R.AverageTransforms([])


def test_type1_reading_filter_and_chain_on_the_corner_scene():
    """The oracle-side composition used as the checker of the type-1 matcher (tests/oracle_lib
    icp_pm_equivalent): the hash filter keeps ~90 % reproducibly, the chain registers the scene."""
    import scenes
    keep = O.pm_keep_mask(100_000, 0.9, 1)
    assert abs(keep.mean() - 0.9) < 5e-3 and np.array_equal(keep, O.pm_keep_mask(100_000, 0.9, 1))
    assert not np.array_equal(keep, O.pm_keep_mask(100_000, 0.9, 2))
    assert O.pm_keep_mask(10, 1.0).all() and not O.pm_keep_mask(10, 0.0).any()
    src, tgt, GT = scenes.corner_pair()
    o = O.icp_pm_equivalent(src.astype(np.float32), tgt.astype(np.float32))
    dt, dr = scenes.se3_error(GT, o["result"])
    assert o["rc"] == 1 and o["ok"] and dt < 0.03 and dr < 5e-3
    assert 0.6 < o["icp_fast_score"] <= o["score"] < 1.0        # raw reference is denser than the filtered one
    assert o["n_target"] == 1024 and 4300 < o["n_source"] < 4700
