"""Committed golden fixtures (tests/golden/config1.npz, made by tests/golden/make_golden.py from
the CPU oracle).  CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them."""
import os

import numpy as np
import pytest

import oracle_lib as O
import scenes
import staticmapping_b200 as smb

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1.npz"))


def test_fixture_inputs_are_the_seeded_scene():
    src, tgt, GT = scenes.corner_pair()
    assert np.array_equal(G["src"], src.astype(np.float32)) and np.array_equal(G["tgt"], tgt.astype(np.float32))
    assert np.allclose(G["GT"], GT)


def test_oracle_reproduces_golden_icp_and_normals():
    tp, tn = O.calculate_normals(G["tgt"].astype(np.float64))
    assert np.array_equal(tp, G["target_points"]) and np.array_equal(tn, G["target_normals"])
    r = O.icp_fast_align(G["src"].astype(np.float64), tp, tn)
    assert r["iterations"] == int(G["icp_iterations"])
    assert np.allclose(r["result"], G["icp_result"], atol=1e-12) and abs(r["score"] - float(G["icp_score"])) < 1e-12
    ids, d2 = O.knn1(tp - tp.mean(0), G["knn_query"], epsilon=3.16)
    assert np.array_equal(ids, G["knn_ids"]) and np.array_equal(d2, G["knn_d2"])


def test_oracle_reproduces_golden_ndt_and_ndt_gicp():
    n = O.ndt_align(G["src"], G["tgt"])
    assert n["iterations"] == int(G["ndt_iterations"]) and np.allclose(n["result"], G["ndt_result"], atol=1e-7)
    g = O.ndt_gicp_align(G["src"], G["tgt"])
    assert np.allclose(g["result"], G["ng_result"], atol=1e-6)
    assert [g["n_source_filtered"], g["n_target_filtered"]] == list(G["ng_counts"][:2])


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    t = smb.CalculateNormals(G["tgt"].astype(np.float64))
    assert np.array_equal(t.points, G["target_points"]) and np.array_equal(t.normals, G["target_normals"])
    ids, d2 = smb.knn1(t.points - t.points.mean(0), G["knn_query"], epsilon=3.16)
    assert np.array_equal(ids, G["knn_ids"]) and np.array_equal(d2, G["knn_d2"])
    m = smb.IcpFast()
    m.SetInputSource(smb.EigenCloud(G["src"].astype(np.float64)))
    m.SetInputTarget(t)
    ok, res = m.Align(np.eye(4))
    dt, dr = scenes.se3_error(G["icp_result"], res)
    assert dt <= 1e-4 and dr <= 1e-4 and m.GetAlignInfo()["iterations"] == int(G["icp_iterations"])
    n = smb.Ndt()
    n.SetInputSource(smb.InnerCloud(G["src"])); n.SetInputTarget(smb.InnerCloud(G["tgt"]))
    ok, res = n.Align(np.eye(4))
    dt, dr = scenes.se3_error(G["ndt_result"], res)
    assert dt <= 1e-4 and dr <= 1e-4 and abs(n.GetFitnessScore() - float(G["ndt_fitness"])) < 1e-9
    g = smb.NdtWithGicp()
    g.SetInputSource(smb.InnerCloud(G["src"])); g.SetInputTarget(smb.InnerCloud(G["tgt"]))
    ok, res = g.Align(np.eye(4))
    dt, dr = scenes.se3_error(G["ng_result"], res)
    assert dt <= 1e-4 and dr <= 1e-4 and abs(g.GetFitnessScore() - float(G["ng_score"])) < 1e-9
    assert g.GetAlignInfo()["aux"][2:] == [float(v) for v in G["ng_counts"][:2]]


# ---- section 8(f) rows: tests/golden/rows_f.npz (tests/golden/make_golden_rows_f.py) -------------------
F = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rows_f.npz"))


def test_oracle_reproduces_golden_rows_f():
    rc, comp = O.motion_compensation(F["scan"], F["delta"])
    assert rc == 0 and np.array_equal(comp, F["compensated"])
    m, vox = O.voxel_grid_filter(F["cloud"], float(F["voxel_size"]))
    assert m == F["voxels"].shape[0] and np.array_equal(vox, F["voxels"])
    src, tgt, _ = scenes.corner_pair()
    pm = O.icp_pm_equivalent(src.astype(np.float32), tgt.astype(np.float32))
    assert [pm["n_source"], pm["n_target"], pm["iterations"], pm["score_kept"]] == list(F["pm_counts"])
    assert np.allclose(pm["result"], F["pm_result"], atol=1e-12) and abs(pm["score"] - float(F["pm_score"])) < 1e-12


@pytest.mark.gpu
def test_gpu_reproduces_golden_rows_f():
    comp = smb.MotionCompensation(F["scan"], F["delta"])
    want = F["compensated"]
    d = np.abs(comp[:, :3].astype(np.float64) - want[:, :3].astype(np.float64))
    assert np.all(d <= np.spacing(np.abs(want[:, :3])).astype(np.float64)) and np.array_equal(comp[:, 3:], want[:, 3:])
    vox = smb.VoxelGridFilter(F["cloud"], float(F["voxel_size"]))
    assert np.array_equal(vox.view(np.uint32), F["voxels"].view(np.uint32))
    src, tgt, _ = scenes.corner_pair()
    m = smb.IcpUsingPointMatcher()
    m.SetInputSource(smb.InnerCloud(src.astype(np.float32)))
    m.SetInputTarget(smb.InnerCloud(tgt.astype(np.float32)))
    ok, res = m.Align(np.eye(4))
    info = m.GetAlignInfo()
    dt, dr = scenes.se3_error(F["pm_result"], res)
    assert ok == bool(F["pm_ok"]) and dt <= 1e-4 and dr <= 1e-4
    assert [int(info["aux"][2]), int(info["aux"][3]), info["iterations"], int(info["aux"][1])] == list(F["pm_counts"])
    assert abs(m.GetFitnessScore() - float(F["pm_score"])) < 1e-9 and abs(info["aux"][0] - float(F["pm_icp_fast_score"])) < 1e-9
