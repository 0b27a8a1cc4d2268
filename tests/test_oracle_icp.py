"""CPU tests of the oracle itself (the reference ships no tests for registrators/, so the
restatement is pinned against analytic known answers, brute force and scipy — SURVEY 8c)."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import oracle_lib as O
import scenes
from staticmapping_b200 import synth


def test_quantile_index_uses_float_ratio():
    # icp_fast.cc:86 with dist_outlier_ratio a float (icp_fast.h:59): 120000 -> 83999
    assert O.quantile_index(120000, 0.7) == 83999
    assert O.quantile_index(10, 0.7) == 6
    assert O.quantile_index(5000, 0.7) == 3499


@pytest.mark.parametrize("n,leaves", [(5000, 1024), (120000, 21696), (500000, 106784)])
def test_build_normals_leaf_counts(n, leaves):
    # cloud_types.cc:105-144 recursion: leaves of <= 7 points
    rng = np.random.default_rng(1)
    p, nrm = O.calculate_normals(rng.uniform(1.0, 50.0, size=(n, 3)))
    assert p.shape[0] == leaves
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-12)


def test_normals_of_planes_are_axis_aligned():
    _, tgt, _ = scenes.corner_pair()
    p, n = O.calculate_normals(tgt)
    dom = np.abs(n).max(axis=1)
    assert np.mean(dom > 0.99) > 0.9          # three axis-aligned planes
    # unconstrained LS plane n.p = 1: normal points away from the origin side
    assert np.all(np.einsum("ij,ij->i", n, p) > 0)


def test_tie_modes_agree_without_ties():
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(3000, 3)) * 10 + 30
    p0, n0 = O.calculate_normals(pts, tie_mode=0)
    p1, n1 = O.calculate_normals(pts, tie_mode=1)
    k0 = np.lexsort(p0.T); k1 = np.lexsort(p1.T)
    assert np.allclose(p0[k0], p1[k1], atol=1e-12) and np.allclose(n0[k0], n1[k1], atol=1e-9)
    T = rng.normal(size=(5000, 3)); Q = rng.normal(size=(2000, 3))
    a = O.knn1(T, Q, epsilon=3.16, tie_mode=0); b = O.knn1(T, Q, epsilon=3.16, tie_mode=1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("nt", [1, 7, 8, 9, 17, 1000, 20000])
def test_exact_knn_matches_brute_force_and_scipy(nt):
    rng = np.random.default_rng(nt)
    T = rng.normal(size=(nt, 3)); Q = rng.normal(size=(1500, 3)) * 1.5
    ids, d2 = O.knn1(T, Q, epsilon=0.0)
    ib, db = O.knn1_brute(T, Q)
    assert np.array_equal(ids, ib) and np.array_equal(d2, db)
    dd, ii = cKDTree(T).query(Q)
    assert np.array_equal(ids, ii.astype(np.int32))
    assert np.allclose(np.sqrt(d2), dd, rtol=1e-13)


def test_approximate_knn_bound():
    # libnabo epsilon semantics: returned distance <= (1 + eps) * true distance
    rng = np.random.default_rng(4)
    T = rng.normal(size=(20000, 3)); Q = rng.normal(size=(5000, 3))
    _, d2a = O.knn1(T, Q, epsilon=3.16)
    _, d2e = O.knn1_brute(T, Q)
    assert np.all(np.sqrt(d2a) <= (1 + 3.16) * np.sqrt(d2e) * (1 + 1e-12))
    assert np.mean(d2a != d2e) > 0.05        # and it really is approximate


def test_solve6_paths():
    rng = np.random.default_rng(0)
    F = rng.normal(size=(6, 100)); A = F @ F.T; b = rng.normal(size=6)
    x, path = O.solve6(A, b)
    assert path == 0 and np.allclose(x, np.linalg.solve(A, b), rtol=1e-10)
    F[3:] = 0.0                               # rank 3 -> reduced min-norm branch (icp_fast.cc:215-232)
    A = F @ F.T; b = A @ rng.normal(size=6)
    x, path = O.solve6(A, b)
    assert path in (1, 2)
    assert np.allclose(x, np.linalg.lstsq(A, b, rcond=None)[0], atol=1e-9)


def test_config1_pose_vs_ground_truth():
    src, tgt, GT = scenes.corner_pair()
    tp, tn = O.calculate_normals(tgt)
    r = O.icp_fast_align(src, tp, tn)
    assert r["rc"] == 1 and 4 <= r["iterations"] <= 100
    dt, dr = scenes.se3_error(GT, r["result"])
    assert dt <= 0.02 and dr <= np.deg2rad(0.2)


def test_identical_clouds_give_identity():
    _, tgt, _ = scenes.corner_pair()
    tp, tn = O.calculate_normals(tgt)
    r = O.icp_fast_align(tp, tp, tn, max_iteration=10)
    assert np.allclose(r["result"], np.eye(4), atol=1e-9) and abs(r["score"] - 1.0) < 1e-12


def test_guess_composition_and_fixed_iterations():
    src, sub, P = scenes.lidar_pair(pair=1)
    tp, tn = O.calculate_normals(sub)
    g = np.eye(4); g[:3, 3] = [0.1, -0.1, 0.02]
    r = O.icp_fast_align(src, tp, tn, guess=g, max_iteration=15, disable_convergence_check=True)
    assert r["iterations"] == 15
    dt, dr = scenes.se3_error(P, r["result"])
    assert dt < 5e-3 and dr < 1e-3
    assert not np.isnan(tn).any()


def test_transform_point_yaw_pi_known_answer():
    # the one adjacent behaviour the reference pins: builder/data/test/test_cloud_types.cc:184-197
    T = synth.se3_from_rpy_t(0.0, 0.0, np.pi, (0, 0, 0))
    p = synth.apply_se3(T, np.array([[10.0, 0.0, 30.0]]))[0]
    assert abs(p[0] + 10.0) < 1e-6 and abs(p[1]) < 1e-6 and abs(p[2] - 30.0) < 1e-6
