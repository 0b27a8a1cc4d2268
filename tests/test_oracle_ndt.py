"""CPU tests of the NDT oracle (restatement of registrators/pclomp/ as configured by
registrators/ndt.cc): known constants, leaf statistics against numpy, derivative
self-consistency, and the reference's clamped-step behaviour."""
import numpy as np

import oracle_lib as O
import scenes


def test_gauss_constants_known_answer():
    # outlier_ratio 0.55, resolution 1 -> d1, d2, d3 = -2.217225, 0.433123, 0.597837
    c1, c2 = 10.0 * (1 - 0.55), 0.55
    d3 = -np.log(c2); d1 = -np.log(c1 + c2) - d3
    d2 = -2 * np.log((-np.log(c1 * np.exp(-0.5) + c2) - d3) / d1)
    assert abs(d1 + 2.217225) < 1e-6 and abs(d2 - 0.433123) < 1e-6 and abs(d3 - 0.597837) < 1e-6
    # and the oracle uses them: one point exactly at a voxel mean scores -d1 per neighbour
    rng = np.random.default_rng(0)
    tgt = (rng.normal(size=(400, 3)) * 0.1 + np.array([0.5, 0.5, 0.5])).astype(np.float32)
    v = O.ndt_voxels(tgt)
    assert len(v["idx"]) == 1 and v["searchable"][0] == 1
    src = v["mean"].astype(np.float32)
    score, g, H, nb = O.ndt_derivatives(src, tgt, np.zeros(6))
    assert nb == 1.0 and abs(score - (-d1)) < 1e-5


def test_leaf_statistics_against_numpy():
    rng = np.random.default_rng(1)
    tgt = (rng.normal(size=(5000, 3)) * np.array([3.0, 2.0, 0.3])).astype(np.float32)
    v = O.ndt_voxels(tgt, 1.0)
    keys = np.floor(tgt.astype(np.float64)).astype(int)
    assert len(v["idx"]) == len(np.unique(keys, axis=0))
    assert v["n"][v["n"] > 0].sum() + np.sum(v["n"] < 0) * 0 <= 5000
    # pick the fullest voxel and recompute: cov = (I + sum xx^T - 2 s m^T)/n + m m^T, *(n-1)/n
    k = int(np.argmax(v["n"]))
    sel = None
    for key in np.unique(keys, axis=0):
        m = np.all(keys == key, axis=1)
        if m.sum() == v["n"][k] and np.allclose(tgt[m].astype(np.float64).mean(0), v["mean"][k], atol=1e-12):
            sel = m
            break
    assert sel is not None
    x = tgt[sel].astype(np.float64)
    n = x.shape[0]
    s = x.sum(0); mean = s / n
    cov = (np.eye(3) + x.T @ x - 2 * np.outer(s, mean)) / n + np.outer(mean, mean)
    cov *= (n - 1.0) / n
    w, V = np.linalg.eigh(cov)
    if w[0] < 0.01 * w[2]:
        w[0] = 0.01 * w[2]
        if w[1] < 0.01 * w[2]:
            w[1] = 0.01 * w[2]
        cov = V @ np.diag(w) @ np.linalg.inv(V)
    assert np.allclose(v["icov"][k], np.linalg.inv(cov), rtol=1e-9, atol=1e-9)
    # leaves below 6 points are not searchable
    assert np.all(v["searchable"][(v["n"] >= 0) & (v["n"] < 6)] == 0)


def test_gradient_and_hessian_match_finite_differences():
    # one voxel, every source point well inside the search radius: the neighbour sets do not
    # change under small pose changes, so the score is smooth and finite differences apply
    rng = np.random.default_rng(3)
    tgt = (rng.normal(size=(600, 3)) * np.array([0.12, 0.08, 0.03]) + 0.5).astype(np.float32)
    assert len(O.ndt_voxels(tgt)["idx"]) == 1
    src = (rng.normal(size=(200, 3)) * 0.15 + 0.5).astype(np.float32)
    p0 = np.array([0.02, -0.03, 0.01, 0.04, -0.03, 0.06])
    score, g, H, nb = O.ndt_derivatives(src, tgt, p0)
    assert nb == 1.0
    h = 2e-3
    for k in range(6):
        pp = p0.copy(); pp[k] += h
        pm = p0.copy(); pm[k] -= h
        sp, gp, _, _ = O.ndt_derivatives(src, tgt, pp)
        sm, gm, _, _ = O.ndt_derivatives(src, tgt, pm)
        fd = (sp - sm) / (2 * h)
        # score = sum(-d1 e), gradient accumulates d1 d2 e x'^T S^-1 J = d(score)/dp (eq. 6.12)
        assert abs(g[k] - fd) <= 2e-2 * max(1.0, abs(fd)), (k, g[k], fd)
        fdh = (gp - gm) / (2 * h)
        assert np.allclose(H[k], fdh, rtol=5e-2, atol=5e-2 * np.abs(H).max()), (k, H[k], fdh)


def test_align_uses_clamped_steps_and_recovers_pose_coarsely():
    src, sub, P = scenes.lidar_pair(pair=2)
    o = O.ndt_align(src.astype(np.float32), sub.astype(np.float32))
    assert o["rc"] == 1 and 1 <= o["iterations"] <= 37
    dt, dr = scenes.se3_error(P, o["result"])
    assert dt < 0.15 and dr < 0.05          # min step 0.05: NDT is the coarse matcher
    assert 0.0 < o["fitness"] < 0.1 and 3.0 < o["mean_neighbors"] < 5.0


def test_invalid_leaf_still_scores():
    # >= 6 collinear-in-a-plane points: eigenvalue test passes or fails, but the centroid stays
    # searchable either way; with icov = 0 the term is exactly -d1 (score) and zero gradient
    t = np.zeros((8, 3), np.float32) + np.array([0.5, 0.5, 0.5], np.float32)   # 8 identical points
    v = O.ndt_voxels(t)
    assert v["searchable"][0] == 1
    score, g, H, nb = O.ndt_derivatives(t[:1], t, np.zeros(6))
    assert nb == 1.0
