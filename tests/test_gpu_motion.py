"""GPU parity of MotionCompensation (builder/map_builder.cc:232-257) against the CPU oracle.
The device evaluates the same double-precision expression tree; sin() may differ from glibc's
in the last bit, so the float outputs are compared to within ONE float ulp (tolerance written
below) and must be bit-identical for the overwhelming majority of the points."""
import ctypes as C

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import oracle_lib as O
import staticmapping_b200 as smb
from staticmapping_b200 import _lib

pytestmark = pytest.mark.gpu


def se3(rpy_deg, t):
    T = np.eye(4)
    T[:3, :3] = Rotation.from_euler("xyz", rpy_deg, degrees=True).as_matrix()
    T[:3, 3] = t
    return T


def scan(n, seed=0):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, 5), np.float32)
    pts[:, :3] = rng.uniform(-80, 80, (n, 3))
    pts[:, 3] = rng.uniform(0, 255, n)
    pts[:, 4] = (np.arange(n) / max(n - 1, 1)).astype(np.float32)   # firing order -> factor
    return pts


def assert_within_one_ulp(got, want):
    assert np.array_equal(got[:, 3:], want[:, 3:])
    d = np.abs(got[:, :3].astype(np.float64) - want[:, :3].astype(np.float64))
    ulp = np.spacing(np.abs(want[:, :3]).astype(np.float32)).astype(np.float64)
    assert np.all(d <= ulp), float(np.max(d / ulp))
    assert np.mean(got[:, :3] == want[:, :3]) > 0.999


@pytest.mark.parametrize("delta", [se3((0.4, -0.3, 2.0), (0.9, -0.1, 0.02)),      # a typical 0.1 s of motion
                                   se3((20, -35, 170), (3.0, -2.0, 1.0)),        # large rotation (d > 0 still)
                                   se3((0, 0, 0), (0.5, 0.0, 0.0)),              # pure translation: lerp branch
                                   se3((180, 0, 0), (0, 0, 0))])                 # w = 0 quaternion
def test_matches_oracle_full_scan(delta):
    pts = scan(120_000, seed=1)
    rc, want = O.motion_compensation(pts, delta)
    assert rc == 0
    got = smb.MotionCompensation(pts, delta)
    assert_within_one_ulp(got, want)


def test_negative_dot_branch():
    # rotation by more than pi about z: Quaternion(Matrix3) returns w < 0 for none of Eigen's
    # branches, so force d < 0 through the (i, j, k) branch: trace <= 0
    delta = se3((0, 0, 179.0), (0, 0, 0)) @ se3((0, 170.0, 0), (0, 0, 0))
    pts = scan(10_000, seed=2)
    rc, want = O.motion_compensation(pts, delta)
    got = smb.MotionCompensation(pts, delta)
    assert rc == 0
    assert_within_one_ulp(got, want)


def test_check_failure_on_bad_factor_and_empty_cloud():
    pts = scan(1000, seed=3)
    pts[17, 4] = 1.5
    with pytest.raises(smb.CheckFailure):
        smb.MotionCompensation(pts, np.eye(4))
    out = smb.MotionCompensation(np.zeros((0, 5), np.float32), np.eye(4))
    assert out.shape == (0, 5)


def test_strided_points_through_the_c_abi():
    # std::vector<InnerPointType> has stride 20; a caller with a wider record passes its stride and
    # the bytes between records come back unchanged
    lib = _lib.lib()
    n = 5000
    pts = scan(n, seed=4)
    wide = np.full((n, 8), -7.0, np.float32)
    wide[:, :5] = pts
    out = np.zeros_like(wide)
    delta = se3((1, 2, 3), (0.1, 0.2, 0.3))
    d = np.asfortranarray(delta)
    rc = lib.sm_motion_compensation(0, wide.ctypes.data, n, 32, d.ctypes.data_as(_lib._DP), out.ctypes.data)
    assert rc == 0
    _, want = O.motion_compensation(pts, delta)
    assert_within_one_ulp(np.ascontiguousarray(out[:, :5]), want)
    assert np.all(out[:, 5:] == -7.0)
    assert lib.sm_motion_compensation(0, wide.ctypes.data, n, 12, d.ctypes.data_as(_lib._DP), out.ctypes.data) == -1
