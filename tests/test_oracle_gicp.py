"""CPU tests of the NdtWithGicp oracle pieces (oracle/gicp_oracle.cc)."""
import numpy as np

import oracle_lib as O
import scenes


def _approx_voxel_grid_reference(pts, leaf):
    """Direct Python transcription of pcl::ApproximateVoxelGrid::applyFilter (sequential)."""
    inv = np.float32(1.0) / np.float32(leaf)
    hist = {}
    out = []
    for p in pts:
        ix, iy, iz = (int(np.floor(np.float32(v) * inv)) for v in p)
        h = (ix * 7171 + iy * 3079 + iz * 4231) & 511
        e = hist.get(h)
        if e is not None and e[0] != (ix, iy, iz):
            out.append(e[1] / np.float32(e[2]))
            e = None
        if e is None:
            e = [(ix, iy, iz), np.zeros(3, np.float32), 0]
            hist[h] = e
        e[1] = e[1] + p.astype(np.float32)
        e[2] += 1
    for h in sorted(hist):
        out.append(hist[h][1] / np.float32(hist[h][2]))
    return np.array(out, np.float32)


def test_approx_voxel_grid_matches_sequential_transcription():
    src, _, _ = scenes.lidar_pair(pair=0)
    pts = src[:3000].astype(np.float32)
    got = O.approx_voxel_grid(pts, 0.2)
    want = _approx_voxel_grid_reference(pts, 0.2)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert 0.3 * len(pts) < len(got) <= len(pts)


def test_gicp_covariances_against_numpy():
    rng = np.random.default_rng(0)
    pts = (rng.normal(size=(400, 3)) * np.array([3.0, 2.0, 0.05])).astype(np.float32)
    cov = O.gicp_covariances(pts, 20, 1e-3)
    d = np.linalg.norm(pts[:, None, :].astype(np.float64) - pts[None, :, :].astype(np.float64), axis=2)
    for i in (0, 17, 399):
        nn = np.argsort(d[i], kind="stable")[:20]
        c = np.cov(pts[nn].astype(np.float64).T, bias=True)
        w, V = np.linalg.eigh(c)                         # ascending
        want = V[:, 2:3] @ V[:, 2:3].T + V[:, 1:2] @ V[:, 1:2].T + 1e-3 * V[:, 0:1] @ V[:, 0:1].T
        assert np.allclose(cov[i], want, atol=1e-6)
        # a thin sheet: the regularised covariance is ~flat in z
        assert cov[i][2, 2] < 0.2


def test_ndt_gicp_refines_pose_and_reports_counts():
    src, sub, P = scenes.lidar_pair(pair=1)
    o = O.ndt_gicp_align(src.astype(np.float32), sub.astype(np.float32))
    assert o["rc"] == 1 and 0 < o["n_source_filtered"] < len(src) and 0 < o["n_target_filtered"] < len(sub)
    assert o["ndt_score"] <= 1.0 and 1 <= o["gicp_iterations"] <= 35 and o["bfgs_evaluations"] > 0
    dt, dr = scenes.se3_error(P, o["result"])
    assert dt < 0.02 and dr < 2e-3
    assert 0.9 < o["score"] <= 1.0            # exp(-mean squared NN distance)


def test_ndt_gate_rejects_bad_alignment():
    # target far away from the source: NDT fitness > 1 -> Align returns false, result = guess,
    # score = exp(-10) (ndt_gicp.cc:104-108)
    src, sub, _ = scenes.lidar_pair(pair=0)
    far = (sub + np.array([500.0, 0.0, 0.0])).astype(np.float32)
    g = np.eye(4); g[0, 3] = 0.25
    o = O.ndt_gicp_align(src.astype(np.float32), far, guess=g)
    assert o["rc"] == 0 and np.array_equal(o["result"], g)
    assert abs(o["score"] - np.exp(-10.0)) < 1e-15
