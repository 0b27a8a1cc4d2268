"""CPU-side checks: the C-ABI library loads and exports every symbol include/sm_b200.h
declares (no compute without a GPU), the host mirror's option/XML logic, sharding."""
import ctypes
import os
import re

import numpy as np
import pytest

import staticmapping_b200 as smb
from staticmapping_b200 import _lib, parallel, registrators, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for header in ("sm_b200.h", "sm_b200_debug.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"typedef[^;]*\(\s*\*\s*\w+\s*\)[^;]*;", "", text)       # function-pointer typedefs are not exports
        names |= set(re.findall(r"\b(sm_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sm_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"


def test_no_cpu_fallback_without_device():
    lib = _lib.lib()
    if lib.sm_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        smb.IcpFast()
    with pytest.raises(RuntimeError):
        smb.knn1(np.zeros((4, 3)), np.zeros((2, 3)))
    with pytest.raises(RuntimeError):
        smb.CalculateNormals(np.ones((10, 3)))


def test_type_enum_matches_reference_xml_values():
    # registrator::Type, interface.h:41-50 / config XML type="6"
    assert int(smb.Type.kIcpPM) == 1 and int(smb.Type.kNdtWithGicp) == 3
    assert int(smb.Type.kNdt) == 5 and int(smb.Type.kFastIcp) == 6


def test_xml_param_parsing():
    xml = ('<registrator_options type="6"><param name="max_iteration"> 30 </param>'
           '<param name="dist_outlier_ratio">0.7</param></registrator_options>')
    assert registrators._params_from_node(xml) == [("max_iteration", "30"), ("dist_outlier_ratio", "0.7")]
    assert registrators._params_from_node({"a": 1}) == [("a", "1")]
    assert registrators._params_from_node(None) == []


def test_eigen_cloud_validation():
    c = smb.EigenCloud.FromPointCloud(np.ones((5, 3), dtype=np.float32))
    assert c.points.dtype == np.float64 and not c.HasNormals()
    with pytest.raises(ValueError):
        smb.EigenCloud(np.ones((5, 2)))
    with pytest.raises(ValueError):
        smb.EigenCloud(np.ones((5, 3)), np.ones((4, 3)))


def test_create_matcher_deprecated_types():
    with pytest.raises(smb.CheckFailure):
        smb.CreateMatcher(smb.MatcherOptions(type=smb.Type.kLibicp))
    assert smb.CreateMatcher(smb.MatcherOptions(type=smb.Type.kNoType)) is None


def test_synth_is_deterministic_and_well_formed():
    sc = synth.make_scene(0)
    a = synth.lidar_scan(sc, (1.0, 0.0, 0.0), seed=3, n_beams=16, n_az=200)
    b = synth.lidar_scan(sc, (1.0, 0.0, 0.0), seed=3, n_beams=16, n_az=200)
    assert a.shape == (3200, 3) and a.dtype == np.float32 and np.array_equal(a, b)
    assert np.isfinite(a).all() and np.linalg.norm(a, axis=1).max() < 81.0
    assert not np.any(a[:, 1] == 0.0)          # no ray exactly in the y = 0 plane
    P = synth.perturbation(5)
    assert np.allclose(P[:3, :3] @ P[:3, :3].T, np.eye(3), atol=1e-12)


def test_shard_pairs_partition():
    for n, w in [(512, 8), (2048, 8), (10, 4), (3, 8), (0, 2)]:
        allp = sum((parallel.shard_pairs(n, r, w) for r in range(w)), [])
        assert allp == list(range(n))
        sizes = [len(parallel.shard_pairs(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_pose_record_roundtrip():
    T = synth.se3_from_rpy_t(0.1, -0.2, 0.3, (1, 2, 3))
    rec = parallel.pack_poses([T, np.eye(4)], [0.5, 1.0])
    assert rec.shape == (2, 17) and rec[0, 12] == 1.0 and rec[0, 13] == 2.0   # column-major
    Ts, s = parallel.unpack_poses(rec)
    assert np.array_equal(Ts[0], T) and s == [0.5, 1.0]
