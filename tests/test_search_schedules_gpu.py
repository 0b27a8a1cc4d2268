"""GPU parity of the k-NN launch shape used with many alignments in flight (`knn_queries_per_cta` > 256:
lockstep root visits, then the lanes of every warp pull the parked far visits, knn_smem.cuh `knn_batch_cta`)
against the CPU oracle, through the test hook sm_debug_knn1_batched (include/sm_b200_debug.h) and through
IcpFast::Align: index sets and squared distances bit-exact, alignments bit-identical to the
one-query-per-thread shape.  Reference call sites: registrators/icp_fast.cc:177-178 (knn), :486-493."""
import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair0():
    return scenes.full_size_pair(0)


@pytest.mark.parametrize("qpc", [512, 768, 1024, 4096])
@pytest.mark.parametrize("eps", [0.0, 3.16])
@pytest.mark.parametrize("nt,nq", [(9, 100), (1000, 5000), (106784, 120000)])
def test_knn_batched_schedule_bit_exact(nt, nq, eps, qpc):
    # the launch shape of batched alignments: lockstep root visits, then the lanes of a warp pull the parked
    # searches (eps = 0 parks nearly every query and makes the far visits long)
    rng = np.random.default_rng(nt * 11 + nq)
    T = rng.normal(size=(nt, 3)) * np.array([20.0, 10.0, 2.0])
    Q = rng.normal(size=(nq, 3)) * np.array([22.0, 11.0, 2.5])
    ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
    ids_g, d2_g = smb.knn1(T, Q, epsilon=eps, queries_per_cta=qpc)
    assert np.array_equal(ids_g, ids_o)
    assert np.array_equal(d2_g, d2_o)


def test_knn_batched_schedule_lidar_scene_and_ties():
    src, sub, _ = scenes.lidar_pair(pair=0)
    ids_o, d2_o = O.knn1(sub, src, epsilon=3.16)
    ids_g, d2_g = smb.knn1(sub, src, epsilon=3.16, queries_per_cta=1024)
    assert np.array_equal(ids_g, ids_o) and np.array_equal(d2_g, d2_o)
    rng = np.random.default_rng(5)
    base = rng.normal(size=(500, 3))
    T = np.concatenate([base, base, base[:100]])        # duplicates: ties resolved by visit order
    Q = np.concatenate([base[:300], rng.normal(size=(700, 3))])
    for eps in (0.0, 3.16):
        ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
        ids_g, d2_g = smb.knn1(T, Q, epsilon=eps, queries_per_cta=512)
        assert np.array_equal(ids_g, ids_o) and np.array_equal(d2_g, d2_o)


@pytest.mark.parametrize("qpc", [512, 1024, 2048])
def test_align_with_batched_search_is_bit_identical(qpc):
    src, sub, _ = scenes.lidar_pair(pair=1)
    tgt = smb.CalculateNormals(sub)
    results = []
    for q in (0, qpc):
        m = smb.IcpFast(0)
        m.InitWithXml({"knn_queries_per_cta": q})
        m.SetInputSource(smb.EigenCloud(src)); m.SetInputTarget(smb.EigenCloud(tgt.points, tgt.normals))
        ok, res = m.Align(np.eye(4))
        results.append((ok, res.copy(), m.GetFitnessScore(), m.GetAlignInfo()["iterations"]))
    assert results[0][0] == results[1][0] and results[0][3] == results[1][3]
    assert np.array_equal(results[0][1], results[1][1]) and results[0][2] == results[1][2]


def test_config2_batched_search_full_size(pair0):
    # the launch shape bench.py uses with many alignments in flight (1 024 queries per CTA: lockstep root
    # visits + warp-pulled far visits): the same index sets, and an alignment bit-identical to the default one
    src, sub, P = pair0
    tp, tn = O.calculate_normals(sub)
    ids_o, d2_o = O.knn1(tp, src, epsilon=3.16)
    ids_g, d2_g = smb.knn1(tp, src, epsilon=3.16, queries_per_cta=1024)
    assert np.array_equal(ids_g, ids_o) and np.array_equal(d2_g, d2_o)
    res = []
    for qpc in (0, 1024):
        m = smb.IcpFast()
        m.InitWithXml({"max_iteration": 30, "disable_convergence_check": 1, "knn_queries_per_cta": qpc})
        m.SetInputSource(smb.EigenCloud(src)); m.SetInputTarget(smb.EigenCloud(tp, tn))
        ok, r = m.Align(np.eye(4))
        assert ok
        res.append((r.copy(), m.GetFitnessScore()))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]


def test_signed_zero_coordinates_split_like_the_oracle():
    # -0.0 and +0.0 are EQUAL coordinates for the reference's comparator (cloud_types.cc:58-66, libnabo alike) and for
    # the oracle's total order (coordinate, index); the sort keys of the tree build canonicalise the sign of zero, so
    # a median that falls inside a group of zeros splits it by index on both sides
    rng = np.random.default_rng(21)
    n = 6000
    x = rng.uniform(-1.0, 1.0, size=n)
    zero = rng.random(n) < 0.5
    x[zero] = np.where(rng.random(int(zero.sum())) < 0.5, -0.0, 0.0)
    T = np.stack([x, rng.normal(size=n) * 0.05, rng.normal(size=n) * 0.05], axis=1)
    assert np.signbit(T[:, 0][T[:, 0] == 0.0]).any() and (~np.signbit(T[:, 0][T[:, 0] == 0.0])).any()
    Q = np.stack([rng.uniform(-0.2, 0.2, size=3000), rng.normal(size=3000) * 0.05, rng.normal(size=3000) * 0.05], axis=1)
    for eps in (0.0, 3.16):
        ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
        for qpc in (0, 1024):
            ids_g, d2_g = smb.knn1(T, Q, epsilon=eps, queries_per_cta=qpc)
            assert np.array_equal(ids_g, ids_o) and np.array_equal(d2_g, d2_o)


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
def test_adversarial_clouds(name, T):
    # the clouds of tests/test_search_formulation_model.py (ties in every split, degenerate extents, both zeros, twelve
    # orders of magnitude between the axes) through the kernels, both launch shapes
    rng = np.random.default_rng(len(name))
    step = max(1, len(T) // 150)
    Q = np.concatenate([T[::step] + rng.normal(size=(len(T[::step]), 3)) * 0.3, rng.normal(size=(100, 3)) * np.abs(T).max()])
    for eps in (0.0, 3.16):
        ids_o, d2_o = O.knn1(T, Q, epsilon=eps)
        for qpc in (0, 512):
            ids_g, d2_g = smb.knn1(T, Q, epsilon=eps, queries_per_cta=qpc)
            assert np.array_equal(ids_g, ids_o) and np.array_equal(d2_g, d2_o)
