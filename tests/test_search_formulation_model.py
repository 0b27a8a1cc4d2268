"""The search FORMULATION of the CUDA kernels (implicit heap, float lower-bound mask, deepest-first far visits,
push-time / pop-time stack tests, tournament bucket scan, warp-pulled parked searches) restated on the CPU
(tests/gpu_knn_model.py) against libnabo's plain recursion (tests/pyref.py) and the C++ oracle: the same index
sets and the same squared distances bit for bit, and the same number of bucket visits per query."""
import numpy as np
import pytest

import oracle_lib as O
import pyref
from gpu_knn_model import GpuKnnModel
from staticmapping_b200 import synth


@pytest.mark.parametrize("eps", [0.0, 0.5, 3.16])
@pytest.mark.parametrize("nt,bucket", [(1, 8), (8, 8), (9, 8), (17, 8), (600, 8), (3000, 8), (1500, 3), (1500, 5)])
def test_iterative_formulation_equals_the_recursion(nt, bucket, eps):
    rng = np.random.default_rng(17 * nt + bucket)
    T = rng.normal(size=(nt, 3)) * np.array([20.0, 10.0, 2.0])
    Q = rng.normal(size=(400, 3)) * np.array([22.0, 11.0, 2.5])
    vr, vm = [], []
    ids_r, d2_r = pyref.PyNabo(T, bucket).knn1(Q, eps, visits=vr)
    ids_m, d2_m = GpuKnnModel(T, bucket).knn1(Q, eps, visits=vm)
    assert np.array_equal(ids_m, ids_r) and np.array_equal(d2_m, d2_r)
    assert vm == vr                                   # not one bucket more or less than the recursion scans
    ids_o, d2_o = O.knn1(T, Q, epsilon=eps, bucket_size=bucket)
    assert np.array_equal(ids_m, ids_o) and np.array_equal(d2_m, d2_o)


@pytest.mark.parametrize("lanes", [1, 8, 32])
def test_warp_pulled_schedule_equals_the_recursion(lanes):
    # which lane runs a parked query, and when, does not change what the query sees
    rng = np.random.default_rng(5)
    T = rng.uniform(-30, 30, size=(4000, 3)) * np.array([1.0, 1.0, 0.05])
    Q = rng.uniform(-30, 30, size=(600, 3)) * np.array([1.0, 1.0, 0.05]) + np.array([0.0, 0.0, 0.4])
    m = GpuKnnModel(T)
    for eps in (0.0, 3.16):
        ids_r, d2_r = pyref.PyNabo(T).knn1(Q, eps)
        ids_b, d2_b = m.knn1_batched(Q, eps, lanes=lanes)
        assert np.array_equal(ids_b, ids_r) and np.array_equal(d2_b, d2_r)


def test_ties_and_duplicates():
    rng = np.random.default_rng(3)
    base = rng.normal(size=(300, 3))
    T = np.concatenate([base, base, base[:50]])
    Q = np.concatenate([base[:200], rng.normal(size=(200, 3))])
    m = GpuKnnModel(T)
    for eps in (0.0, 3.16):
        ids_r, d2_r = pyref.PyNabo(T).knn1(Q, eps)
        for ids_m, d2_m in (m.knn1(Q, eps), m.knn1_batched(Q, eps)):
            assert np.array_equal(ids_m, ids_r) and np.array_equal(d2_m, d2_r)


def test_lidar_geometry_and_the_float_lower_bound_mask_is_a_superset():
    scene = synth.make_scene(0)
    scan = synth.lidar_scan(scene, (0.0, 0.0, 0.0), seed=3).astype(np.float64)
    T, Q = scan[::30], scan[7::200] + np.array([0.3, -0.2, 0.05])
    m = GpuKnnModel(T)
    ids_r, d2_r = pyref.PyNabo(T).knn1(Q, 3.16)
    ids_m, d2_m = m.knn1(Q, 3.16)
    assert np.array_equal(ids_m, ids_r) and np.array_equal(d2_m, d2_r)
    # the mask built from round-down float bounds contains every level the exact test passes
    me2 = 4.16 * 4.16
    for q in Q[:200]:
        q = tuple(float(v) for v in q)
        st = {"head": np.inf, "best": -1, "visits": 0}
        hp1, ll, mask = m._root_visit(q, me2, st)
        for a in range(ll):
            n = (hp1 >> (ll - a)) - 1
            off = q[m.dim[n]] - m.cut[n]
            if (off * off) * me2 < st["head"]:
                assert mask & (1 << a)


def _adversarial_clouds():
    rng = np.random.default_rng(0)
    yield "grid", np.stack(np.meshgrid(np.arange(12.0), np.arange(9.0), np.arange(5.0)), -1).reshape(-1, 3)
    yield "line", np.stack([np.linspace(-3, 3, 700), np.zeros(700), np.zeros(700)], 1)
    yield "identical", np.tile([[1.5, -2.0, 0.25]], (100, 1))
    yield "signed-zeros", np.stack([rng.integers(-20, 20, 900).astype(float), rng.integers(-20, 20, 900).astype(float),
                                    np.where(rng.random(900) < 0.5, -0.0, 0.0)], 1)
    yield "huge-range", rng.normal(size=(1500, 3)) * np.array([1e6, 1e-6, 1.0])
    yield "quantised", np.round(rng.normal(size=(3000, 3)) * 4) / 4
    yield "two-clusters", np.concatenate([rng.normal(size=(800, 3)) * 0.01, rng.normal(size=(800, 3)) * 0.01 + 100])


@pytest.mark.parametrize("name,T", list(_adversarial_clouds()), ids=[n for n, _ in _adversarial_clouds()])
def test_adversarial_clouds_all_four_statements_agree(name, T):
    # grids and quantised coordinates (ties in every split), degenerate extents (argmax of all-zero extents is axis 0),
    # both zeros, twelve orders of magnitude between the axes: oracle, Python recursion, kernel formulation, warp-pulled
    rng = np.random.default_rng(len(name))
    step = max(1, len(T) // 150)
    Q = np.concatenate([T[::step] + rng.normal(size=(len(T[::step]), 3)) * 0.3, rng.normal(size=(100, 3)) * np.abs(T).max()])
    m = GpuKnnModel(T)
    for eps in (0.0, 3.16):
        ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
        for ids, d2 in (pyref.PyNabo(T).knn1(Q, eps), m.knn1(Q, eps), m.knn1_batched(Q, eps)):
            assert np.array_equal(ids, ids_o) and np.array_equal(d2, d2_o)
