"""GPU parity of VoxelGrid::Filter (pre_processors/filter_voxel_grid.cc:37-78) against the oracle:
bit-identical voxel means, in the engine's ascending-voxel order."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

pytestmark = pytest.mark.gpu


def as5(xyz, seed=0):
    rng = np.random.default_rng(seed)
    p = np.zeros((xyz.shape[0], 5), np.float32)
    p[:, :3] = xyz
    p[:, 3] = rng.uniform(0, 255, xyz.shape[0])
    p[:, 4] = rng.uniform(0, 1, xyz.shape[0])
    return p


@pytest.mark.parametrize("voxel", [0.1, 0.2, 1.0])
def test_submap_bit_exact(voxel):
    _, sub, _ = scenes.lidar_pair(pair=1)          # 60 k-point submap of the synthetic scene
    pts = as5(sub.astype(np.float32), seed=1)
    m, want = O.voxel_grid_filter(pts, voxel)
    got = smb.VoxelGridFilter(pts, voxel)
    assert got.shape == (m, 5)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))      # bit for bit


def test_full_size_dense_voxels_and_ties():
    rng = np.random.default_rng(9)
    n = 500_000
    xyz = rng.uniform(-4, 4, (n, 3)).astype(np.float32)                   # ~1000 points per 0.8 m voxel
    xyz[:1000, 0] = np.float32(0.4)                                       # exactly on a voxel boundary (x / 0.8 = 0.5)
    xyz[1000:2000, 0] = np.float32(-0.4)
    pts = as5(xyz, seed=2)
    m, want = O.voxel_grid_filter(pts, 0.8)
    got = smb.VoxelGridFilter(pts, 0.8)
    assert got.shape[0] == m and np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_edge_cases():
    one = np.array([[1.0, -2.0, 3.0, 7.0, 0.9]], np.float32)
    assert np.array_equal(smb.VoxelGridFilter(one, 0.5), np.array([[1.0, -2.0, 3.0, 7.0, 0.0]], np.float32))
    assert smb.VoxelGridFilter(np.zeros((0, 5), np.float32), 0.5).shape == (0, 5)
    with pytest.raises(smb.CheckFailure):
        smb.VoxelGridFilter(one, 0.0)                                     # ConfigsValid
    bad = one.copy(); bad[0, 1] = np.nan                                  # a lone non-finite point is dropped
    assert smb.VoxelGridFilter(bad, 0.5).shape == (0, 5)
    same = np.tile(one, (5000, 1))                                        # everything in one voxel
    out = smb.VoxelGridFilter(same, 0.5)
    assert out.shape == (1, 5) and np.array_equal(out[0, :4], one[0, :4])


def test_non_finite_points_are_dropped_and_far_clouds_work():
    # one bad lidar return must not stop the mapper: NaN / inf points are dropped, the rest is the
    # filter of the finite points; voxel keys are offsets from the cloud's own minimum, so a cloud
    # 150 km from the origin with 0.1 m voxels (index ~1.5e6 > 2^20) is fine
    rng = np.random.default_rng(3)
    pts = np.zeros((20000, 5), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, (20000, 3)) + np.array([150_000.0, -90_000.0, 30.0])
    pts[:, 3] = rng.uniform(0, 255, 20000)
    m, want = O.voxel_grid_filter(pts, 0.1)
    got = smb.VoxelGridFilter(pts, 0.1)
    assert got.shape[0] == m and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    dirty = pts.copy()
    dirty[7, 0] = np.nan; dirty[11, 2] = np.inf; dirty[13, 1] = -np.inf
    clean = np.delete(pts, [7, 11, 13], axis=0)
    m2, want2 = O.voxel_grid_filter(clean, 0.1)
    got2 = smb.VoxelGridFilter(dirty, 0.1)
    assert got2.shape[0] == m2 and np.array_equal(got2.view(np.uint32), want2.view(np.uint32))
    wide = pts[:2].copy(); wide[1, 0] += 3.0e6                            # spans > 2^21 voxels of 0.1 m
    with pytest.raises(smb.CheckFailure):
        smb.VoxelGridFilter(wide, 0.1)


def test_reference_unit_test_vectors():
    """pre_processors/test/test_filter_voxel_grid.cc:52-100: 100 / 36 / 9 voxels at 0.1 / 0.2 / 0.4, invalid
    size rejected — the golden values the reference itself holds for this row."""
    pts = scenes.reference_voxel_test_cloud()
    for voxel, want in ((0.1, 100), (0.2, 36), (0.4, 9), (10.0, 1)):
        got = smb.VoxelGridFilter(pts, voxel)
        m, ref = O.voxel_grid_filter(pts, voxel)
        assert got.shape[0] == want == m and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    with pytest.raises(smb.CheckFailure):
        smb.VoxelGridFilter(pts, 0.0)
