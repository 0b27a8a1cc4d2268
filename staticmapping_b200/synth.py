"""Synthetic scan generators for the BASELINE.json configs.

Pure numpy, deterministic per seed, no file I/O.  These only *produce inputs*; nothing
here is on the timed or graded path.

* ``corner_scene_cloud``  - config 1: 3-plane corner, 5 000 points, 1 cm noise.
* ``lidar_scan``          - configs 2-5: 64-beam x 1 875-azimuth ray cast (120 000 float32
                             points in firing order) of ground + two walls + 40 boxes.
* ``submap``              - union of 5 scans at x = 0..4 m sampled to exactly 500 000.
* ``perturbation``        - per-pair SE(3) perturbation, rng(1000 + i).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "se3_from_rpy_t", "apply_se3", "corner_scene_cloud", "corner_ground_truth",
    "make_scene", "lidar_scan", "submap", "perturbation",
]


def se3_from_rpy_t(roll, pitch, yaw, t):
    """4x4 double, R = Rz(yaw) Ry(pitch) Rx(roll)."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = rz @ ry @ rx
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T


def apply_se3(T, pts):
    """pts: (N,3) -> (N,3) double."""
    pts = np.asarray(pts, dtype=np.float64)
    return pts @ T[:3, :3].T + T[:3, 3]


# --------------------------------------------------------------------------- config 1
def corner_scene_cloud(n=5000, seed=1234, sigma=0.01):
    """Points on the corner {z=-1.7, x=+8, y=-6} inside a 20 m box (sensor at origin,
    so no plane passes through the origin - CalculateNormals' n.p=1 model needs that)."""
    rng = np.random.default_rng(seed)
    which = rng.integers(0, 3, size=n)
    u = rng.uniform(-10.0, 10.0, size=n)
    v = rng.uniform(-10.0, 10.0, size=n)
    noise = rng.normal(0.0, sigma, size=n)
    pts = np.empty((n, 3), dtype=np.float64)
    g = which == 0  # ground z = -1.7
    pts[g] = np.stack([u[g] * 0.8, v[g] * 0.6 + 0.0, -1.7 + noise[g]], axis=1)
    w = which == 1  # wall x = +8
    pts[w] = np.stack([8.0 + noise[w], v[w] * 0.6, (u[w] + 10.0) * 0.25 - 1.7], axis=1)
    s = which == 2  # wall y = -6
    pts[s] = np.stack([u[s] * 0.8, -6.0 + noise[s], (v[s] + 10.0) * 0.25 - 1.7], axis=1)
    return pts.astype(np.float32)  # clouds enter the reference as float (InnerPointType)


def corner_ground_truth():
    return se3_from_rpy_t(np.deg2rad(1.0), np.deg2rad(-2.0), np.deg2rad(3.0),
                          (0.30, -0.20, 0.10))


# ----------------------------------------------------------------------- configs 2..5
def make_scene(seed=0, n_boxes=40):
    """Axis-aligned boxes standing on the ground z=-1.73 between walls y=+-10."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-40.0, 45.0, size=n_boxes)
    cy = rng.uniform(-9.0, 9.0, size=n_boxes)
    sx = rng.uniform(0.5, 4.0, size=n_boxes)
    sy = rng.uniform(0.5, 4.0, size=n_boxes)
    h = rng.uniform(0.5, 3.5, size=n_boxes)
    # keep a corridor around the sensor track (x in [-1, 6], |y| < 1.5) free
    blocked = (np.abs(cy) - sy / 2 < 1.5) & (cx + sx / 2 > -1.0) & (cx - sx / 2 < 6.0)
    cy = np.where(blocked, np.sign(cy + 1e-9) * (1.6 + sy / 2 + 0.5), cy)
    lo = np.stack([cx - sx / 2, cy - sy / 2, np.full(n_boxes, -1.73)], axis=1)
    hi = np.stack([cx + sx / 2, cy + sy / 2, -1.73 + h], axis=1)
    return {"ground_z": -1.73, "wall_y": 10.0, "lo": lo, "hi": hi}


def _cast(scene, origin, dirs, max_range):
    """Range of the first hit for each ray (inf if none)."""
    o = np.asarray(origin, dtype=np.float64)
    d = dirs
    best = np.full(d.shape[0], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (scene["ground_z"] - o[2]) / d[:, 2]
        best = np.where((t > 0) & (t < best), t, best)
        for sgn in (-1.0, 1.0):
            t = (sgn * scene["wall_y"] - o[1]) / d[:, 1]
            best = np.where((t > 0) & (t < best), t, best)
        inv = 1.0 / d
        t0 = (scene["lo"][None, :, :] - o[None, None, :]) * inv[:, None, :]
        t1 = (scene["hi"][None, :, :] - o[None, None, :]) * inv[:, None, :]
        tn = np.nanmax(np.minimum(t0, t1), axis=2)
        tf = np.nanmin(np.maximum(t0, t1), axis=2)
        hit = (tf >= tn) & (tf > 0)
        tb = np.where(hit, np.where(tn > 0, tn, tf), np.inf).min(axis=1)
        best = np.minimum(best, tb)
    best[best > max_range] = np.inf
    return best


def lidar_scan(scene, sensor_xyz=(0.0, 0.0, 0.0), seed=0, n_beams=64, n_az=1875,
               sigma=0.02, max_range=80.0):
    """(n_beams*n_az, 3) float32 points in the SENSOR frame, firing order
    (azimuth-major, beam-minor).  Rays without a return are re-drawn."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    # half-step offset: no ray has an exactly zero x or y component, so no leaf of
    # CalculateNormals lies exactly in a plane through the sensor origin (singular M)
    az = (np.arange(n_az) + 0.5) * (2.0 * np.pi / n_az)
    A, E = np.meshgrid(az, elev, indexing="ij")
    A = A.ravel().copy()
    E = E.ravel().copy()
    n = A.size
    rng_out = np.full(n, np.inf)
    todo = np.arange(n)
    for _ in range(64):
        d = np.stack([np.cos(E[todo]) * np.cos(A[todo]), np.cos(E[todo]) * np.sin(A[todo]),
                      np.sin(E[todo])], axis=1)
        r = _cast(scene, sensor_xyz, d, max_range)
        rng_out[todo] = r
        miss = ~np.isfinite(r)
        if not miss.any():
            break
        todo = todo[miss]
        A[todo] = rng.uniform(0.0, 2.0 * np.pi, size=todo.size)
        E[todo] = rng.choice(elev, size=todo.size)
    else:
        raise RuntimeError("ray re-draw did not terminate")
    rng_out = rng_out + rng.normal(0.0, sigma, size=n)
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=1)
    return (d * rng_out[:, None]).astype(np.float32)


def submap(scene, seed=0, n_points=500_000, n_scans=5, n_beams=64, n_az=1875):
    """Union of n_scans scans taken at x = 0,1,..  expressed in the frame of scan 0,
    sampled (rng(seed+1)) to exactly n_points float32 points."""
    parts = []
    for k in range(n_scans):
        pos = (float(k), 0.0, 0.0)
        s = lidar_scan(scene, pos, seed=seed * 131 + k, n_beams=n_beams, n_az=n_az)
        parts.append(s.astype(np.float64) + np.asarray(pos))
    allp = np.concatenate(parts, axis=0)
    rng = np.random.default_rng(seed + 1)
    if allp.shape[0] >= n_points:
        sel = np.sort(rng.choice(allp.shape[0], size=n_points, replace=False))
    else:
        sel = np.sort(rng.choice(allp.shape[0], size=n_points, replace=True))
    return allp[sel].astype(np.float32)


def perturbation(i):
    """Per-pair source perturbation of SURVEY 8d: rng(1000+i)."""
    rng = np.random.default_rng(1000 + i)
    t = rng.uniform(-1.0, 1.0, size=3) * np.array([0.5, 0.5, 0.1])
    yaw = np.deg2rad(rng.uniform(-2.0, 2.0))
    roll, pitch = np.deg2rad(rng.uniform(-0.5, 0.5, size=2))
    return se3_from_rpy_t(roll, pitch, yaw, t)
