"""The M2DP restatement (oracle/m2dp_oracle.cc) against a plain numpy transcription of
descriptor/m2dp.cc with an independent linear algebra (numpy eigh / svd), and its structural properties."""
import numpy as np

import oracle_lib as O
import scenes


def _numpy_m2dp(pts, r=0.1, max_distance=100.0, t=16, p=4, q=16):
    x = pts.astype(np.float32)
    mean = x.astype(np.float64).mean(axis=0).astype(np.float32)
    d = x - mean
    C = (d.astype(np.float64).T @ d.astype(np.float64) / np.float64(np.float32(len(x) - 1))).astype(np.float32)
    w, V = np.linalg.eigh(C.astype(np.float64))
    V = V[:, ::-1]
    E = np.zeros((3, 3))
    for j in range(2):
        v = V[:, j]
        E[:, j] = v * (1.0 if v[np.argmax(np.abs(v))] >= 0 else -1.0)
    E = E.astype(np.float32)
    E[:, 2] = np.cross(E[:, 0], E[:, 1])
    P = (d @ E).astype(np.float32)
    P = P[np.sqrt((P * P).sum(axis=1, dtype=np.float32)) <= max_distance]
    l = int(np.ceil(np.sqrt(max_distance / r)))
    A = np.zeros((p * q, l * t))
    for pi in range(p):
        for qi in range(q):
            th, ph = pi * np.pi / p, qi * (np.pi / 2) / q
            m = np.array([np.cos(th) * np.cos(ph), np.cos(th) * np.sin(ph), np.sin(th)]).astype(np.float32)
            xa = (np.array([1, 0, 0], np.float32) - np.abs(m[0]) * m).astype(np.float32)
            ya = np.cross(m, xa).astype(np.float32)
            u = np.abs(P @ xa).astype(np.float32); v = np.abs(P @ ya).astype(np.float32)
            length = np.sqrt(u * u + v * v).astype(np.float64)
            ang = np.arctan2(v, u).astype(np.float64)
            li = np.minimum(np.floor(np.sqrt(length / r)).astype(int), l - 1)
            ti = np.minimum(np.floor(ang / (2 * np.pi / t)).astype(int), t - 1)
            np.add.at(A[pi * q + qi], li * t + ti, 1)
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    u1, v1 = U[:, 0], Vt[0]
    sgn = 1.0 if u1[np.argmax(np.abs(u1))] >= 0 else -1.0
    return np.concatenate([sgn * u1, sgn * v1]).astype(np.float32), A


def test_oracle_against_numpy_transcription():
    src, sub, P = scenes.lidar_pair(pair=0)
    pts = sub.astype(np.float32)
    d, A, axes = O.m2dp(pts, with_matrix=True)
    dn, An = _numpy_m2dp(pts)
    assert A.sum() == An.sum() == 64 * len(pts)
    assert np.abs(A - An).sum() / 2 <= 2e-3 * A.sum()        # float sums in another order move a few border points
    assert O.m2dp_match(d, dn) > 0.99999
    assert np.abs(d - dn).max() < 2e-3


def test_only_the_first_quadrant_is_filled():
    # `(p.transpose() * axis).norm()` is an absolute value (m2dp.cc:103-104): angles stay in [0, pi/2]
    src, sub, P = scenes.lidar_pair(pair=1)
    d, A, axes = O.m2dp(sub.astype(np.float32), with_matrix=True)
    t = 16
    used = np.unique(np.nonzero(A)[1] % t)
    assert set(used.tolist()) <= {0, 1, 2, 3, 4}
    assert abs(np.linalg.norm(d[:64]) - 1) < 1e-5 and abs(np.linalg.norm(d[64:]) - 1) < 1e-5


def test_match_score():
    a = np.linspace(0, 1, 50).astype(np.float32)
    assert abs(O.m2dp_match(a, a) - 1.0) < 1e-6
    assert abs(O.m2dp_match(a, -a) - 1.0) < 1e-6            # fabs(score)
    assert O.m2dp_match(a[:5], a[:5]) == -1.0               # fewer than 10 rows (m2dp.cc:157)
    assert O.m2dp(np.zeros((0, 3), np.float32)) is None     # empty cloud: setInputCloud returns false
