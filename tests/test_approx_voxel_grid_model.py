"""pcl::ApproximateVoxelGrid::applyFilter (registrators/ndt_gicp.cc:59-70) is ONE sequential pass over a 512-slot hash
history.  The CUDA path (staticmapping_b200/csrc/gicp.cu approx_*) computes it in parallel: stable sort by slot ->
runs of equal voxel inside a slot -> float centroid per run in original order -> runs ordered by the index of the point
that evicts them, never-evicted runs last in slot order.  This restates that formulation with numpy and checks that it
emits the sequential pass's output, count and bits, including on clouds built to collide in the hash."""
import numpy as np
import pytest

import oracle_lib as O
import scenes


def sequential(pts, leaf):
    """direct transcription of the PCL loop"""
    inv = np.float32(1.0) / np.float32(leaf)
    hist, out = {}, []
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
    return np.array(out, np.float32).reshape(-1, 3)


def parallel_formulation(pts, leaf):
    """the kernels' steps: approx_key -> stable radix sort -> run heads -> approx_run (centroid + emission key) -> sort
    by emission key -> approx_emit"""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    n = pts.shape[0]
    inv = np.float32(1.0) / np.float32(leaf)
    cell = np.floor(pts * inv).astype(np.int64)
    slot = (cell[:, 0] * 7171 + cell[:, 1] * 3079 + cell[:, 2] * 4231) & 511
    order = np.argsort(slot, kind="stable")
    skey = slot[order]
    scell = cell[order]
    head = np.ones(n, dtype=bool)
    head[1:] = (skey[1:] != skey[:-1]) | np.any(scell[1:] != scell[:-1], axis=1)
    run_start = np.r_[np.flatnonzero(head), n]
    centroids, emit_key = [], []
    for s0, s1 in zip(run_start[:-1], run_start[1:]):
        c = np.zeros(3, np.float32)
        for k in range(s0, s1):                       # float sum in original order
            c = c + pts[order[k]]
        centroids.append(c / np.float32(s1 - s0))
        evicted = s1 < n and skey[s1] == skey[s0]
        emit_key.append(int(order[s1]) if evicted else n + int(skey[s0]))
    emit = np.argsort(np.array(emit_key), kind="stable")
    return np.array(centroids, np.float32).reshape(-1, 3)[emit]


def _colliding_cloud(rng, n):
    # few distinct voxels, many of them sharing hash slots, visited in an interleaved order: many evictions
    cells = rng.integers(-40, 40, size=(60, 3))
    pick = rng.integers(0, 60, size=n)
    return ((cells[pick] + rng.random((n, 3))) * 0.2).astype(np.float32)


@pytest.mark.parametrize("case", ["lidar", "colliding", "one-voxel", "negative-coordinates"])
def test_parallel_formulation_emits_the_sequential_output(case):
    rng = np.random.default_rng(11)
    if case == "lidar":
        src, _, _ = scenes.lidar_pair(pair=0)
        pts = src[:4000].astype(np.float32)
    elif case == "colliding":
        pts = _colliding_cloud(rng, 5000)
    elif case == "one-voxel":
        pts = (rng.random((500, 3)) * 0.19 + 1.0).astype(np.float32)
    else:
        pts = (rng.normal(size=(3000, 3)) * 3.0 - 5.0).astype(np.float32)
    want = sequential(pts, 0.2)
    got = parallel_formulation(pts, 0.2)
    assert got.shape == want.shape
    assert np.array_equal(got, want)                                  # count, order and bits
    assert np.array_equal(O.approx_voxel_grid(pts, 0.2), want)        # and the C++ oracle is the sequential pass
