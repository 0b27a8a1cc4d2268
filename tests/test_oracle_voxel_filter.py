"""Known-answer checks of the VoxelGrid restatement (oracle/motion_oracle.cc; reference
pre_processors/filter_voxel_grid.cc:37-78).  No GPU."""
import numpy as np

import oracle_lib as O
import scenes


def cloud(n, seed=0, span=20.0):
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 5), np.float32)
    p[:, :3] = rng.uniform(-span, span, (n, 3))
    p[:, 3] = rng.uniform(0, 255, n)
    p[:, 4] = rng.uniform(0, 1, n)
    return p


def test_hand_computed_voxels():
    pts = np.array([[0.04, 0.0, 0.0, 10, 0.5],      # lround(0.4) = 0
                    [0.05, 0.0, 0.0, 20, 0.5],      # lround(0.5) = 1 : half rounds away from zero
                    [-0.05, 0.0, 0.0, 30, 0.5],     # lround(-0.5) = -1
                    [0.14, 0.0, 0.0, 40, 0.5],      # 1
                    [0.0, 0.26, -0.31, 50, 0.5]], np.float32)
    m, out = O.voxel_grid_filter(pts, 0.1)
    assert m == 4
    # ascending (ix, iy, iz): (-1,0,0) (0,0,0) (0,3,-3) (1,0,0)
    assert np.allclose(out[0], [-0.05, 0, 0, 30, 0]) and np.allclose(out[1], [0.04, 0, 0, 10, 0])
    assert np.allclose(out[2], [0.0, 0.26, -0.31, 50, 0])
    want = (np.float64(np.float32(0.05)) + np.float64(np.float32(0.14))) / 2
    assert out[3, 0] == np.float32(want) and out[3, 3] == np.float32(30.0) and out[3, 4] == 0.0


def test_literal_unordered_map_order_is_the_same_set():
    pts = cloud(50_000, seed=3)
    m0, a = O.voxel_grid_filter(pts, 0.5, order_mode=0)
    m1, b = O.voxel_grid_filter(pts, 0.5, order_mode=1)
    assert m0 == m1 and 0 < m0 < pts.shape[0]
    key = lambda x: np.lexsort((x[:, 2], x[:, 1], x[:, 0]))
    assert np.array_equal(a[key(a)], b[key(b)])             # identical values, different order
    assert not np.array_equal(a, b)


def test_means_and_invalid_size():
    pts = cloud(20_000, seed=5, span=3.0)
    m, out = O.voxel_grid_filter(pts, 0.25)
    idx = np.array([np.rint(np.float64(pts[:, d] / np.float32(0.25))) for d in range(3)]).T   # no exact .5 ties here
    uniq, inv = np.unique(idx, axis=0, return_inverse=True)
    assert m == uniq.shape[0]
    for v in (0, m // 2, m - 1):
        sel = pts[inv.ravel() == v]
        assert np.allclose(out[v, :4], sel[:, :4].astype(np.float64).mean(axis=0), atol=1e-5)
    assert O.voxel_grid_filter(pts, 0.0)[0] == -1


def test_pinned_by_the_reference_unit_test():
    """BOOST_AUTO_TEST_CASE(Filter) / (FilterConfig) of test_filter_voxel_grid.cc: voxel 0.1 keeps all 100
    points (:76), 0.2 leaves 36 (:87), 0.4 leaves 9 (:98); voxel_size 0 is an invalid config (:37-41),
    10 a valid one (:44-48).  These are the only golden values the reference holds for this row."""
    pts = scenes.reference_voxel_test_cloud()
    for order_mode in (0, 1):
        assert O.voxel_grid_filter(pts, 0.1, order_mode)[0] == 100
        assert O.voxel_grid_filter(pts, 0.2, order_mode)[0] == 36
        assert O.voxel_grid_filter(pts, 0.4, order_mode)[0] == 9
    assert O.voxel_grid_filter(pts, 0.0)[0] == -1
    assert O.voxel_grid_filter(pts, 10.0)[0] == 1
