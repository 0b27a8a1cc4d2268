"""GPU parity: libsm_b200 (CUDA, through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

pytestmark = pytest.mark.gpu

TOL_T = 1e-4   # metres      (BASELINE.json north_star)
TOL_R = 1e-4   # radians


@pytest.mark.parametrize("eps", [0.0, 3.16])
@pytest.mark.parametrize("nt,nq", [(1, 10), (8, 100), (9, 100), (1000, 5000), (106784, 120000)])
def test_knn_index_sets_bit_exact(nt, nq, eps):
    rng = np.random.default_rng(nt * 7 + nq)
    T = rng.normal(size=(nt, 3)) * np.array([20.0, 10.0, 2.0])
    Q = rng.normal(size=(nq, 3)) * np.array([22.0, 11.0, 2.5])
    ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
    ids_g, d2_g = smb.knn1(T, Q, epsilon=eps)
    assert np.array_equal(ids_g, ids_o)
    assert np.array_equal(d2_g, d2_o)          # bit-exact squared distances


@pytest.mark.parametrize("bucket", [2, 3, 4, 5, 7])
def test_knn_other_bucket_sizes(bucket):
    # libnabo's bucketSize parameter; one padded bucket of the compact layout holds up to 8 points
    rng = np.random.default_rng(100 + bucket)
    T = rng.normal(size=(3000, 3)) * np.array([20.0, 10.0, 2.0])
    Q = rng.normal(size=(4000, 3)) * np.array([22.0, 11.0, 2.5])
    for eps in (0.0, 3.16):
        ids_o, d2_o = O.knn1(T, Q, epsilon=eps, bucket_size=bucket)
        ids_g, d2_g = smb.knn1(T, Q, epsilon=eps, bucket_size=bucket)
        assert np.array_equal(ids_g, ids_o)
        assert np.array_equal(d2_g, d2_o)


def test_knn_bucket_above_8_is_rejected():
    T = np.zeros((100, 3)); Q = np.zeros((10, 3))
    with pytest.raises(RuntimeError):
        smb.knn1(T, Q, bucket_size=16)


@pytest.mark.parametrize("eps", [0.0, 3.16])
def test_knn_deep_tree_top_levels_in_shared_memory(eps):
    # 300 000 points = 16 levels: the top 14 are staged in shared memory, the rest is read from global
    rng = np.random.default_rng(77)
    T = rng.normal(size=(300_000, 3)) * np.array([30.0, 30.0, 3.0])
    Q = rng.normal(size=(30_000, 3)) * np.array([33.0, 33.0, 3.5])
    ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
    ids_g, d2_g = smb.knn1(T, Q, epsilon=eps)
    assert np.array_equal(ids_g, ids_o)
    assert np.array_equal(d2_g, d2_o)


def test_knn_duplicates_and_ties():
    rng = np.random.default_rng(5)
    base = np.round(rng.normal(size=(300, 3)) * 4.0, 1)       # many exact duplicates
    T = np.concatenate([base, base[:100]], axis=0)
    Q = np.round(rng.normal(size=(2000, 3)) * 4.0, 1)
    for eps in (0.0, 3.16):
        ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
        ids_g, d2_g = smb.knn1(T, Q, epsilon=eps)
        assert np.array_equal(d2_g, d2_o)
        assert np.array_equal(ids_g, ids_o)


def _align_both(src, tp, tn, guess=None, **opts):
    m = smb.IcpFast()
    m.InitWithXml({k: v for k, v in opts.items()})
    m.SetInputSource(smb.EigenCloud(src))
    m.SetInputTarget(smb.EigenCloud(tp, tn))
    g = np.eye(4) if guess is None else guess
    ok, res = m.Align(g)
    o = O.icp_fast_align(src, tp, tn, guess=g,
                         max_iteration=int(opts.get("max_iteration", 100)),
                         dist_outlier_ratio=float(opts.get("dist_outlier_ratio", 0.7)),
                         knn_epsilon=float(opts.get("knn_epsilon", 3.16)),
                         disable_convergence_check=bool(int(opts.get("disable_convergence_check", 0))))
    return ok, res, m, o


def test_config1_corner_scene():
    src, tgt, GT = scenes.corner_pair()
    tp, tn = O.calculate_normals(tgt)
    assert tp.shape[0] == 1024
    ok, res, m, o = _align_both(src, tp, tn)
    assert ok and o["rc"] == 1
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)
    info = m.GetAlignInfo()
    assert info["iterations"] == o["iterations"]
    assert abs(m.GetFitnessScore() - o["score"]) < 1e-6
    gt_t, gt_r = scenes.se3_error(GT, res)
    assert gt_t <= 0.02 and gt_r <= np.deg2rad(0.2)


@pytest.mark.parametrize("pair", [0, 1, 2])
def test_lidar_pair_parity(pair):
    src, sub, P = scenes.lidar_pair(pair=pair)
    tp, tn = O.calculate_normals(sub)
    ok, res, m, o = _align_both(src, tp, tn, max_iteration=30)
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)
    assert m.GetAlignInfo()["iterations"] == o["iterations"]
    assert abs(m.GetFitnessScore() - o["score"]) < 1e-6
    # and the alignment is real: the perturbation is recovered (noise-limited)
    gt_t, gt_r = scenes.se3_error(P, res)
    assert gt_t < 5e-3 and gt_r < 1e-3, (gt_t, gt_r)
    assert m.GetAlignInfo()["solve_path"] == 0


def test_nan_normal_in_target_matches_oracle():
    # A leaf lying exactly in a plane through the origin gives a NaN normal in the
    # reference (cloud_types.cc:93, singular M).  IcpFast then sees a NaN normal matrix;
    # the restated Eigen logic (rank 0 -> min-norm -> SVD) yields x = 0.  Same on both sides.
    src, sub, P = scenes.lidar_pair(pair=0)
    tp, tn = O.calculate_normals(sub)
    tn = tn.copy(); tn[5] = np.nan
    ok, res, m, o = _align_both(src, tp, tn, max_iteration=30)
    assert o["rc"] == 1
    assert np.array_equal(np.isnan(res), np.isnan(o["result"]))
    fin = ~np.isnan(res)
    assert np.allclose(res[fin], o["result"][fin], atol=1e-9)
    assert m.GetAlignInfo()["iterations"] == o["iterations"]


def test_fixed_iterations_and_guess():
    src, sub, P = scenes.lidar_pair(pair=3)
    tp, tn = O.calculate_normals(sub)
    guess = np.eye(4); guess[:3, 3] = [0.05, -0.03, 0.01]
    ok, res, m, o = _align_both(src, tp, tn, guess=guess, max_iteration=12,
                                disable_convergence_check=1)
    assert m.GetAlignInfo()["iterations"] == 12 == o["iterations"]
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)


def test_identical_clouds_identity():
    # icp_fast.cc:315-321: b = x = 0, rotation undetermined -> identity
    src, tgt, _ = scenes.corner_pair()
    tp, tn = O.calculate_normals(tgt)
    ok, res, m, o = _align_both(tp.copy(), tp, tn, max_iteration=10)
    assert np.allclose(res, np.eye(4), atol=1e-9)
    assert np.allclose(o["result"], np.eye(4), atol=1e-9)
    assert abs(m.GetFitnessScore() - 1.0) < 1e-9


def test_single_plane_rank_deficient():
    # icp_fast.cc:214-249: one plane -> rank-3 normal matrix -> min-norm branch
    rng = np.random.default_rng(9)
    n = 4000
    tgt = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), np.full(n, -1.7)], axis=1)
    tp, tn = O.calculate_normals(tgt)
    T = np.eye(4); T[:3, 3] = [0.0, 0.0, 0.12]
    src = np.stack([rng.uniform(-8, 8, 3000), rng.uniform(-8, 8, 3000), np.full(3000, -1.7)], axis=1)
    src = src @ T[:3, :3].T - T[:3, 3]
    ok, res, m, o = _align_both(src, tp, tn, max_iteration=8)
    assert m.GetAlignInfo()["solve_path"] in (1, 2)
    dt, dr = scenes.se3_error(o["result"], res)
    assert dt <= TOL_T and dr <= TOL_R, (dt, dr)
    assert abs(res[2, 3] - 0.12) < 1e-6


def test_unknown_option_is_check_failure():
    m = smb.IcpFast()
    with pytest.raises(smb.CheckFailure):
        m.InitWithXml('<registrator_options type="6"><param name="bogus">1</param></registrator_options>')


def test_target_without_normals_is_check_failure():
    m = smb.IcpFast()
    with pytest.raises(smb.CheckFailure):
        m.SetInputTarget(smb.EigenCloud(np.zeros((10, 3)) + 1.0))


def test_align_pairs_and_batch_match_single_aligns():
    # sm_align_pairs (pipelined, async uploads) and sm_align_batch must give what sm_align gives
    data = []
    for pair in range(5):
        src, sub, P = scenes.lidar_pair(pair=pair)
        tp, tn = O.calculate_normals(sub)
        data.append((np.ascontiguousarray(src), np.ascontiguousarray(tp), np.ascontiguousarray(tn)))
    singles = []
    for src, tp, tn in data:
        m = smb.IcpFast()
        m.InitWithXml({"max_iteration": 30})
        m.SetInputSource(smb.EigenCloud(src)); m.SetInputTarget(smb.EigenCloud(tp, tn))
        ok, res = m.Align(np.eye(4))
        singles.append((res, m.GetFitnessScore(), m.GetAlignInfo()["iterations"]))
    ms = [smb.IcpFast() for _ in range(2)]
    for m in ms:
        m.InitWithXml({"max_iteration": 30})
    rcs, res, scores = smb.AlignPairs(ms, [{"source": s, "target": t, "normals": n} for s, t, n in data])
    assert list(rcs) == [1] * 5
    for k in range(5):
        assert np.array_equal(res[k], singles[k][0])          # same kernels, same order: bit-identical
        assert scores[k] == singles[k][1]
    ms3 = []
    for src, tp, tn in data[:3]:
        m = smb.IcpFast()
        m.InitWithXml({"max_iteration": 30})
        m.SetInputSource(smb.EigenCloud(src)); m.SetInputTarget(smb.EigenCloud(tp, tn))
        ms3.append(m)
    oks, resb = smb.AlignBatch(ms3, [np.eye(4)] * 3)
    assert oks == [True] * 3
    for k in range(3):
        assert np.array_equal(resb[k], singles[k][0])


def test_bad_outlier_ratio_is_check_failure():
    # CHECK(quantile >= 0 && quantile <= 1), icp_fast.cc:68
    src, tgt, _ = scenes.corner_pair()
    tp, tn = O.calculate_normals(tgt)
    for bad in ("1.5", "-0.1", "nan"):
        m = smb.IcpFast()
        m.InitWithXml({"dist_outlier_ratio": bad})
        m.SetInputSource(smb.EigenCloud(src)); m.SetInputTarget(smb.EigenCloud(tp, tn))
        with pytest.raises(smb.CheckFailure):
            m.Align(np.eye(4))
