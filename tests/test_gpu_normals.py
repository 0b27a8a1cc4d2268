"""GPU parity of the target preparation (EigenPointCloud::CalculateNormals,
builder/data/cloud_types.cc:73-144,347-368) against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb
from staticmapping_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,expect", [(5, 1), (7, 1), (8, 2), (100, 16), (5000, 1024), (120000, 21696),
                                      (500000, 106784)])
def test_leaf_counts_and_bit_exact_output(n, expect):
    # SURVEY 8c known answers: 5000 -> 1024, 120000 -> 21696, 500000 -> 106784 leaves
    rng = np.random.default_rng(n)
    pts = rng.uniform(1.0, 100.0, size=(n, 3))
    op, on = O.calculate_normals(pts)
    g = smb.CalculateNormals(pts)
    assert op.shape[0] == expect
    assert g.points.shape[0] == expect
    assert np.array_equal(g.points, op)       # same op order, -fmad=false: bit-identical
    assert np.array_equal(g.normals, on)


def test_corner_scene_normals():
    _, tgt, _ = scenes.corner_pair()
    op, on = O.calculate_normals(tgt)
    g = smb.CalculateNormals(tgt)
    assert np.array_equal(g.points, op) and np.array_equal(g.normals, on)
    assert np.allclose(np.linalg.norm(g.normals, axis=1), 1.0, atol=1e-12)


def test_lidar_submap_normals_with_float_ties():
    # float32 lidar data has many exactly equal coordinates -> exercises the
    # (coordinate, index) total order of the median split
    scene = synth.make_scene(3)
    sub = synth.submap(scene, seed=3, n_points=40000, n_scans=3, n_beams=16, n_az=600).astype(np.float64)
    op, on = O.calculate_normals(sub)
    g = smb.CalculateNormals(sub)
    assert g.points.shape == op.shape
    assert np.array_equal(g.points, op) and np.array_equal(g.normals, on)


def test_degenerate_leaves_are_dropped():
    # collinear points: covariance rank 1 -> rank + 1 < 3 -> leaf skipped (cloud_types.cc:89)
    t = np.linspace(1.0, 50.0, 64)
    pts = np.stack([t, 2.0 * t, 0.5 * t + 1.0], axis=1)
    op, on = O.calculate_normals(pts)
    g = smb.CalculateNormals(pts)
    assert op.shape[0] == g.points.shape[0] == 0
