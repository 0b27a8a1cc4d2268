"""Second, independent restatement of the ICP half of the path in pure Python / numpy.

TEST INFRASTRUCTURE ONLY.  The C++ oracle (oracle/sm_oracle.cc) is the checker of the CUDA path; the
reference ships no tests or golden vectors for registrators/ and cannot be compiled in this image
(SURVEY.md 8c), so the oracle itself is "parity unpinned".  This module narrows that gap from the other
side: the same algorithms written a second time, in another language, directly from the reference
sources / the published libnabo algorithm — NOT from the oracle's code — and compared with the oracle
in tests/test_oracle_vs_python_restatement.py.  A transcription slip in either restatement shows up as a
disagreement.  Python floats are IEEE doubles without contraction, so everything that is comparison /
index logic plus a fixed order of +, -, * is expected to agree BIT FOR BIT (k-NN index sets and squared
distances, leaf partitions); sums over many points and the 6x6 solve go through numpy (other summation
order, LAPACK instead of Eigen) and agree to rounding.

Follows:
  PyNabo            libnabo 1.0.7 nabo/kdtree_cpu.cpp, KDTreeUnbalancedPtInLeavesImplicitBoundsStackOpt
                    (buildNodes, recurseKnn) as called at registrators/icp_fast.cc:466-467 (create) and
                    :177-178 (knn, k = 1, epsilon = 3.16, ALLOW_SELF_MATCH); SURVEY.md Appendix C
  calculate_normals builder/data/cloud_types.cc:41-56 (ArgMax), :73-103 (leaf), :105-144 (BuildNormals),
                    :347-368 (CalculateNormals)
  icp_fast_align    registrators/icp_fast.cc:65-90 (quantile), :100-166 (ErrorElements), :182-202
                    (CrossProduct), :204-254 (solve), :256-324 (ComputePointToPlane), :377-405
                    (CheckConvergence), :456-529 (Align); cloud_types.cc:288-296 (ApplyTransform)
"""
from __future__ import annotations

import math
import sys

import numpy as np

INF = float("inf")


def _arg_max(v):
    """cloud_types.cc:41-56 / libnabo argMax: first index of the largest STRICTLY positive entry, else 0."""
    max_val, max_idx = 0.0, 0
    for i in range(len(v)):
        if v[i] > max_val:
            max_val, max_idx = v[i], i
    return max_idx


def _median_split(points, idx, cut_dim):
    """std::nth_element(first, first + leftCount, last, CompareDim(cut_dim)) as a SET partition:
    rightCount = count / 2, leftCount = count - rightCount.  Equal coordinates are ordered by the
    original index (the reference leaves that to nth_element's internals; with distinct coordinates
    the partition is unique)."""
    count = len(idx)
    right_count = count // 2
    left_count = count - right_count
    order = sorted(idx, key=lambda i: (points[i][cut_dim], i))
    return order[:left_count], order[left_count:]


# ------------------------------------------------------------------------------------ libnabo
class PyNabo:
    """NNS::create(cloud, 3, KDTREE_LINEAR_HEAP) with the default bucket size 8, then knn(k = 1)."""

    def __init__(self, cloud, bucket_size=8):
        self.pts = [tuple(float(x) for x in p) for p in np.asarray(cloud, dtype=np.float64)]
        self.bucket_size = bucket_size
        self.nodes = []          # inner: (dim, cut_val, right_child) / leaf: (3, [indices], None)
        n = len(self.pts)
        if n == 0:
            return
        a = np.asarray(cloud, dtype=np.float64)
        min_values = [float(a[:, r].min()) for r in range(3)]
        max_values = [float(a[:, r].max()) for r in range(3)]
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
        self._build(list(range(n)), min_values, max_values)

    def _build(self, idx, min_values, max_values):
        pos = len(self.nodes)
        if len(idx) <= self.bucket_size:
            # bucket entries in the order of the index array; canonical order here = ascending index
            self.nodes.append((3, sorted(idx), None))
            return pos
        cut_dim = _arg_max([max_values[r] - min_values[r] for r in range(3)])
        left, right = _median_split(self.pts, idx, cut_dim)
        cut_val = self.pts[right[0]][cut_dim]          # coordinate of *(first + leftCount)
        self.nodes.append(None)
        left_max = list(max_values); left_max[cut_dim] = cut_val
        right_min = list(min_values); right_min[cut_dim] = cut_val
        self._build(left, min_values, left_max)        # left child = pos + 1
        rc = self._build(right, right_min, max_values)
        self.nodes[pos] = (cut_dim, cut_val, rc)
        return pos

    def knn1(self, query, epsilon=3.16, visits=None):
        """-> (ids int32[nq] with -1 = InvalidIndex, squared distances f64[nq] with +inf = InvalidValue)"""
        q_all = np.asarray(query, dtype=np.float64)
        ids = np.full(q_all.shape[0], -1, dtype=np.int32)
        d2 = np.full(q_all.shape[0], INF, dtype=np.float64)
        if not self.nodes:
            return ids, d2
        max_error2 = (1.0 + epsilon) * (1.0 + epsilon)
        for j in range(q_all.shape[0]):
            q = (float(q_all[j, 0]), float(q_all[j, 1]), float(q_all[j, 2]))
            self._head, self._head_idx, self._visits = INF, -1, 0
            self._recurse(q, 0, 0.0, [0.0, 0.0, 0.0], max_error2)
            ids[j], d2[j] = self._head_idx, self._head
            if visits is not None:
                visits.append(self._visits)
        return ids, d2

    def _recurse(self, q, n, rd, off, max_error2):
        dim, a, right_child = self.nodes[n]
        if dim == 3:                                   # leaf: entries in order, strict '<' replaces the head
            self._visits += 1
            for index in a:
                p = self.pts[index]
                dist = 0.0
                for r in range(3):
                    diff = q[r] - p[r]
                    dist += diff * diff
                if dist < self._head:                  # dist <= maxRadius2 (= inf) always holds
                    self._head, self._head_idx = dist, index
            return
        old_off = off[dim]
        new_off = q[dim] - a
        if new_off > 0.0:
            self._recurse(q, right_child, rd, off, max_error2)
            rd += -old_off * old_off + new_off * new_off
            if rd <= INF and rd * max_error2 < self._head:
                off[dim] = new_off
                self._recurse(q, n + 1, rd, off, max_error2)
                off[dim] = old_off
        else:
            self._recurse(q, n + 1, rd, off, max_error2)
            rd += -old_off * old_off + new_off * new_off
            if rd <= INF and rd * max_error2 < self._head:
                off[dim] = new_off
                self._recurse(q, right_child, rd, off, max_error2)
                off[dim] = old_off


# ------------------------------------------------------------------------------------ target prep
def calculate_normals(points, leaf_size=7):
    """EigenPointCloud::CalculateNormals: -> (points (M,3), normals (M,3), leaves) with one row per kept
    leaf, rows ordered by the leaf's smallest original index, `leaves` = the index sets of ALL leaves."""
    pts = np.asarray(points, dtype=np.float64)
    rows = [tuple(float(x) for x in p) for p in pts]
    kept, leaves = [], []

    def leaf(idx):
        leaves.append(sorted(idx))
        d = pts[sorted(idx)].T                                     # 3 x count
        m_wave = d @ d.T
        b_wave = d.sum(axis=1)
        mean = b_wave / d.shape[1]
        nn = d - mean[:, None]
        c = nn @ nn.T
        # fullPivHouseholderQr().rank() + 1 < 3  -> skipped
        if np.linalg.matrix_rank(c) + 1 < 3:
            return
        normal = np.linalg.solve(m_wave, b_wave)                   # M_wave.inverse() * b_wave
        kept.append((min(idx), mean, normal / np.linalg.norm(normal)))

    def build(idx, min_values, max_values):
        if len(idx) <= leaf_size:
            leaf(idx)
            return
        cut_dim = _arg_max([max_values[r] - min_values[r] for r in range(3)])
        left, right = _median_split(rows, idx, cut_dim)
        cut_val = rows[right[0]][cut_dim]
        left_max = list(max_values); left_max[cut_dim] = cut_val
        right_min = list(min_values); right_min[cut_dim] = cut_val
        build(left, min_values, left_max)
        build(right, right_min, max_values)

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    build(list(range(len(rows))), [float(x) for x in pts.min(axis=0)], [float(x) for x in pts.max(axis=0)])
    kept.sort(key=lambda t: t[0])                                  # std::sort(indices_to_keep)
    return (np.array([k[1] for k in kept]).reshape(-1, 3), np.array([k[2] for k in kept]).reshape(-1, 3), leaves)


# ------------------------------------------------------------------------------------ IcpFast
def _apply_transform(T, P):
    """cloud_types.cc:288-296: (T * [P; 1]) rows 0..2; P is (N,3)."""
    # Eigen evaluates every output coefficient as the inner product ((T_i0 x + T_i1 y) + T_i2 z) + T_i3 * 1
    # (no FMA with the reference's flags, -O2 without -march); a BLAS matmul would contract / re-associate
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    return np.stack([((T[i, 0] * x + T[i, 1] * y) + T[i, 2] * z) + T[i, 3] * 1.0 for i in range(3)], axis=1)


def _quantile_limit(d2, ratio_f32):
    """Matches::GetDistsQuantile with dist_outlier_ratio a FLOAT option (icp_fast.h:59)."""
    values = d2[d2 != INF]
    assert values.size > 0
    q = float(np.float32(ratio_f32))
    assert 0.0 <= q <= 1.0
    if q == 1.0:
        return float(values.max())
    qi = int(values.size * q)
    return float(np.partition(values, qi)[qi])


def _solve_possibly_underdetermined(A, b):
    """icp_fast.cc:204-254.  Invertible -> the unique solution (LLT in the reference); otherwise the
    minimum-norm solution (the reference's rank-reduced QR construction and its SVD fallback both
    target it)."""
    # FullPivHouseholderQR::isInvertible(): rank == 6 with threshold eps * 6 on the pivots; numpy's
    # matrix_rank uses S.max() * 6 * eps on the singular values — the same notion of numerical rank
    if np.linalg.matrix_rank(A) == A.shape[0]:
        return np.linalg.solve(A, b)
    return np.linalg.lstsq(A, b, rcond=None)[0]


def _compute_point_to_plane(P, Q, N):
    """icp_fast.cc:256-324 with all weights 1 (ErrorElements compacts the zero weights away)."""
    cross = np.cross(P, N)                                         # reading x normal
    F = np.concatenate([cross, N], axis=1)                         # K x 6
    A = F.T @ F
    dot = np.einsum("ij,ij->i", P - Q, N)
    b = -(F.T @ dot)
    x = _solve_possibly_underdetermined(A, b)
    T = np.eye(4)
    angle = float(np.linalg.norm(x[:3]))
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = x[:3] / angle                                       # .normalized(): 0/0 -> NaN like Eigen < 3.3? see below
    if angle > 0.0 and np.all(np.isfinite(axis)):
        K = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
        T[:3, :3] = np.eye(3) + math.sin(angle) * K + (1.0 - math.cos(angle)) * (K @ K)
    # angle == 0 (identical clouds): the reference's matrix comes out NaN or identity depending on
    # Eigen's normalized(); ":315-321 if result.hasNaN() -> rotation = identity" makes both identity
    T[:3, 3] = x[3:6]
    if not np.all(np.isfinite(T)):
        T[:3, :3] = np.eye(3)
    return T


def _quat_from_matrix(R):
    """Eigen::Quaterniond(Matrix3d) (Shepperd's branches as in Eigen/src/Geometry/Quaternion.h)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0.0:
        s = math.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([w, (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s])
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[1 + i] = 0.5 * s
    s = 0.5 / s
    q[0] = (R[k, j] - R[j, k]) * s
    q[1 + j] = (R[j, i] + R[i, j]) * s
    q[1 + k] = (R[k, i] + R[i, k]) * s
    return q


def _angular_distance(a, b):
    """QuaternionBase::angularDistance (Eigen 3.3): d = a * conj(b); 2 * atan2(|d.vec|, |d.w|)."""
    w = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]
    v = -a[0] * b[1:] + b[0] * a[1:] - np.cross(a[1:], b[1:])
    return 2.0 * math.atan2(float(np.linalg.norm(v)), abs(w))


def _check_convergence(rotations, translations):
    k_smooth = 4
    if len(rotations) <= k_smooth:
        return False
    rot = trans = 0.0
    for i in range(len(rotations) - 1, len(rotations) - 1 - k_smooth, -1):
        rot += abs(_angular_distance(rotations[i], rotations[i - 1]))
        trans += abs(float(np.linalg.norm(translations[i] - translations[i - 1])))
    return rot / k_smooth < 0.001 and trans / k_smooth < 0.01


def icp_fast_align(source, target, target_normals, guess=None, max_iteration=100, dist_outlier_ratio=0.7,
                   disable_convergence_check=False, knn=None, trace=None):
    """IcpFast::Align.  knn(target_centred, P) -> (ids, d2): defaults to PyNabo (pure Python, small clouds)."""
    S = np.asarray(source, dtype=np.float64)
    Qt = np.asarray(target, dtype=np.float64).copy()
    Nq = np.asarray(target_normals, dtype=np.float64)
    guess = np.eye(4) if guess is None else np.asarray(guess, dtype=np.float64)
    target_mean = Qt.sum(axis=0) / Qt.shape[0]
    T_mean = np.eye(4); T_mean[:3, 3] = target_mean
    Qt -= target_mean
    if knn is None:
        tree = PyNabo(Qt)
        knn = lambda _t, P: tree.knn1(P, 3.16)      # noqa: E731
    G0 = np.linalg.inv(T_mean) @ guess
    S0 = _apply_transform(G0, S)
    T_iter = np.eye(4)
    rotations, translations = [np.array([1.0, 0.0, 0.0, 0.0])], [np.zeros(3)]
    it = 0
    while True:
        P = _apply_transform(T_iter, S0)
        ids, d2 = knn(Qt, P)
        limit = _quantile_limit(d2, dist_outlier_ratio)
        keep = np.nonzero((d2 != INF) & (d2 <= limit))[0]
        assert keep.size > 0
        Pk, Qk, Nk = P[keep], Qt[ids[keep]], Nq[ids[keep]]
        T_iter = _compute_point_to_plane(Pk, Qk, Nk) @ T_iter
        it += 1
        rotations.append(_quat_from_matrix(T_iter[:3, :3]))
        translations.append(T_iter[:3, 3].copy())
        if trace is not None:
            trace.append({"limit": limit, "kept": int(keep.size), "T_iter": T_iter.copy()})
        conv = (not disable_convergence_check) and _check_convergence(rotations, translations)
        if conv or it >= max_iteration:
            score = math.exp(-float(np.sqrt(d2[keep]).sum()) / keep.size)
            break
    return {"result": T_mean @ T_iter @ G0, "iterations": it, "score": score, "kept": int(keep.size)}


# ------------------------------------------------------------------------------------ Ndt (pclomp)
# Follows registrators/ndt.cc:29-64 (resolution 1.0, KDTREE neighbour search) and
# registrators/pclomp/ndt_omp_impl.hpp:47-171 (constructor constants, computeTransformation), :180-284
# (computeDerivatives), :288-393 (computeAngleDerivatives), :397-438 (computePointDerivatives, float),
# :483-535 (updateDerivatives), :757-916 (computeStepLengthMT: `interval_converged = (step_max - step_min) > 0`
# is TRUE on entry, so the More-Thuente loop never runs and every step is step_init clamped to
# [transformation_epsilon / 2, step_size]); voxel_grid_covariance_omp_impl.hpp:49-370 (applyFilter) and
# voxel_grid_covariance_omp.h:96-106 (Leaf(): cov_ starts as IDENTITY), :283-297, :471-499 (radiusSearch).
# External: PCL KdTreeFLANN::radiusSearch (float squared distances, strict '<', sorted),
# pcl::Registration::getFitnessScore (mean squared distance to the exact nearest target point).
# Written with numpy float32 arrays for the reference's Matrix<float, ...> math: the ORDER of float
# operations differs from the oracle's scalar code, so agreement is to float rounding (~1e-5 relative).
F32 = np.float32


def ndt_gauss_constants(outlier_ratio=0.55, resolution=1.0):
    c1 = 10.0 * (1.0 - outlier_ratio)
    c2 = outlier_ratio / resolution ** 3
    d3 = -math.log(c2)
    d1 = -math.log(c1 + c2) - d3
    d2 = -2.0 * math.log((-math.log(c1 * math.exp(-0.5) + c2) - d3) / d1)
    return d1, d2, d3


class NdtGrid:
    """VoxelGridCovariance::filter(true) over a float32 cloud."""

    def __init__(self, target_f32, resolution=1.0, min_points=6, eig_mult=0.01):
        t = np.ascontiguousarray(target_f32, dtype=F32)
        inv = F32(1.0) / F32(resolution)
        min_p, max_p = t.min(axis=0), t.max(axis=0)
        min_b = np.floor(min_p * inv).astype(np.int64)
        max_b = np.floor(max_p * inv).astype(np.int64)
        div_b = max_b - min_b + 1
        mul = np.array([1, div_b[0], div_b[0] * div_b[1]], dtype=np.int64)
        ijk = (np.floor(t * inv) - min_b.astype(F32)).astype(np.int64)      # float subtraction, then int cast
        idx = ijk @ mul
        order = np.argsort(idx, kind="stable")                               # std::map: ascending idx; points keep input order
        sidx = idx[order]
        starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
        ends = np.r_[starts[1:], sidx.size]
        self.leaf_idx, self.n, self.mean, self.icov, self.centroid, self.searchable = [], [], [], [], [], []
        for s, e in zip(starts, ends):
            pf = t[order[s:e]]
            x = pf.astype(np.float64)
            n = e - s
            pt_sum = x.sum(axis=0)
            cov = np.eye(3) + x.T @ x                                        # Leaf(): cov_ = Identity, then += x x^T
            centroid = np.cumsum(pf, axis=0, dtype=F32)[-1] / F32(n)         # float accumulation in input order
            mean = pt_sum / n
            icov = np.zeros((3, 3))
            searchable, npts = 0, n
            if n >= min_points:
                searchable = 1                                               # pushed to voxel_centroids_ before the tests
                cov = (cov - 2.0 * np.outer(pt_sum, mean)) / n + np.outer(mean, mean)
                cov *= (n - 1.0) / n
                w, V = np.linalg.eigh(cov)
                if w[0] < 0 or w[1] < 0 or w[2] <= 0:
                    npts = -1
                else:
                    mn = eig_mult * w[2]
                    if w[0] < mn:
                        w[0] = mn
                        if w[1] < mn:
                            w[1] = mn
                        cov = V @ np.diag(w) @ np.linalg.inv(V)
                    icov = np.linalg.inv(cov)
                    if np.isinf(icov).any():
                        npts = -1
            self.leaf_idx.append(int(sidx[s])); self.n.append(npts); self.mean.append(mean)
            self.icov.append(icov); self.centroid.append(centroid); self.searchable.append(searchable)
        self.n = np.array(self.n); self.mean = np.array(self.mean); self.icov = np.array(self.icov)
        self.centroid = np.array(self.centroid, dtype=F32); self.searchable = np.array(self.searchable)
        self.resolution = resolution
        from scipy.spatial import cKDTree
        self._sel = np.flatnonzero(self.searchable == 1)
        self._tree = cKDTree(self.centroid[self._sel].astype(np.float64))

    def radius_search(self, pts_f32):
        """-> (point index, leaf index) pairs, per point sorted by the float squared distance."""
        cand = self._tree.query_ball_point(pts_f32.astype(np.float64), self.resolution * 1.001)
        pi = np.repeat(np.arange(len(cand)), [len(c) for c in cand])
        li = self._sel[np.fromiter((j for c in cand for j in c), dtype=np.int64, count=pi.size)]
        d = pts_f32[pi] - self.centroid[li]                                  # FLANN L2_Simple<float>
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = d2 < F32(self.resolution * self.resolution)
        pi, li, d2 = pi[keep], li[keep], d2[keep]
        o = np.lexsort((d2, pi))
        return pi[o], li[o]


def _ndt_angle_derivatives(p, dt=F32):
    def cs(a):
        return (1.0, 0.0) if abs(a) < 10e-5 else (math.cos(a), math.sin(a))
    cx, sx = cs(p[3]); cy, sy = cs(p[4]); cz, sz = cs(p[5])
    j_ang = np.array([
        [-sx * sz + cx * sy * cz, -sx * cz - cx * sy * sz, -cx * cy],
        [cx * sz + sx * sy * cz, cx * cz - sx * sy * sz, -sx * cy],
        [-sy * cz, sy * sz, cy],
        [sx * cy * cz, -sx * cy * sz, sx * sy],
        [-cx * cy * cz, cx * cy * sz, -cx * sy],
        [-cy * sz, -cy * cz, 0.0],
        [cx * cz - sx * sy * sz, -cx * sz - sx * sy * cz, 0.0],
        [sx * cz + cx * sy * sz, cx * sy * cz - sx * sz, 0.0]]).astype(dt)
    h_ang = np.array([
        [-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, sx * cy],          # a2
        [-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, -cx * cy],         # a3
        [cx * cy * cz, -cx * cy * sz, cx * sy],                               # b2
        [sx * cy * cz, -sx * cy * sz, sx * sy],                               # b3
        [-sx * cz - cx * sy * sz, sx * sz - cx * sy * cz, 0.0],               # c2
        [cx * cz - sx * sy * sz, -sx * sy * cz - cx * sz, 0.0],               # c3
        [-cy * cz, cy * sz, sy],                                              # d1
        [-sx * sy * cz, sx * sy * sz, sx * cy],                               # d2
        [cx * sy * cz, -cx * sy * sz, -cx * cy],                              # d3
        [sy * sz, sy * cz, 0.0],                                              # e1
        [-sx * cy * sz, -sx * cy * cz, 0.0],                                  # e2
        [cx * cy * sz, cx * cy * cz, 0.0],                                    # e3
        [-cy * cz, cy * sz, 0.0],                                             # f1
        [-cx * sz - sx * sy * cz, -cx * cz + sx * sy * sz, 0.0],              # f2
        [-sx * sz + cx * sy * cz, -cx * sy * sz - sx * cz, 0.0]]).astype(dt)   # f3
    return j_ang, h_ang


def ndt_derivatives(grid, source_f32, trans_f32, p, outlier_ratio=0.55, f64=False):
    """computeDerivatives(score_gradient, hessian, trans_cloud, p, true) -> (score, g[6], H[6,6], mean neighbours).
    f64 = False: pclomp (Matrix<float, 4, 6> point math, registrators/pclomp/ndt_omp_impl.hpp);
    f64 = True : stock pcl::NormalDistributionsTransform (the same formulas in double, pcl/registration/impl/ndt.hpp),
                 the matcher NdtWithGicp drives (registrators/ndt_gicp.cc:44-47,82-88)."""
    dt = np.float64 if f64 else F32
    d1, d2, _ = ndt_gauss_constants(outlier_ratio, grid.resolution)
    j_ang, h_ang = _ndt_angle_derivatives(p, dt)
    pi, li = grid.radius_search(trans_f32)
    m = pi.size
    x = source_f32[pi].astype(np.float64)                    # Eigen::Vector3d x(x_pt.x, ...)
    x4 = x.astype(dt)                                         # pclomp: Vector4f x4(x[0], x[1], x[2], 0)
    xj = x4 @ j_ang.T                                         # (m, 8)
    PG = np.zeros((m, 3, 6), dtype=dt)                        # rows 0..2 of the 4x6 (row 3 is zero)
    PG[:, 0, 0] = PG[:, 1, 1] = PG[:, 2, 2] = 1.0
    PG[:, 1, 3] = xj[:, 0]; PG[:, 2, 3] = xj[:, 1]
    PG[:, 0, 4] = xj[:, 2]; PG[:, 1, 4] = xj[:, 3]; PG[:, 2, 4] = xj[:, 4]
    PG[:, 0, 5] = xj[:, 5]; PG[:, 1, 5] = xj[:, 6]; PG[:, 2, 5] = xj[:, 7]
    xh = x4 @ h_ang.T                                         # (m, 15)
    z = np.zeros(m, dtype=dt)
    a = np.stack([z, xh[:, 0], xh[:, 1]], axis=1); b = np.stack([z, xh[:, 2], xh[:, 3]], axis=1)
    c = np.stack([z, xh[:, 4], xh[:, 5]], axis=1); d = xh[:, 6:9]; e = xh[:, 9:12]; f = xh[:, 12:15]
    PH = np.zeros((m, 6, 3, 6), dtype=dt)                     # block i (rows 4i..4i+2 / 3i..3i+2), column j
    PH[:, 3, :, 3] = a; PH[:, 4, :, 3] = b; PH[:, 5, :, 3] = c
    PH[:, 3, :, 4] = b; PH[:, 4, :, 4] = d; PH[:, 5, :, 4] = e
    PH[:, 3, :, 5] = c; PH[:, 4, :, 5] = e; PH[:, 5, :, 5] = f
    xt = (trans_f32[pi].astype(np.float64) - grid.mean[li]).astype(dt)       # x_trans (double) [-> x_trans4 (float)]
    ci = grid.icov[li].astype(dt)                                            # c_inv [.cast<float>()]
    gd2 = dt(d2)
    xc = np.einsum("mi,mij->mj", xt, ci)                                     # x_trans * c_inv
    q = np.einsum("mj,mj->m", xt, xc)
    with np.errstate(over="ignore", invalid="ignore"):
        e_x = np.exp(-gd2 * q * dt(0.5)).astype(dt)
    score_inc = (-d1 * e_x.astype(np.float64)).astype(dt)                    # pclomp: float score_inc = -gauss_d1_ * e
    e_x = gd2 * e_x
    valid = ~((e_x > 1) | (e_x < 0) | np.isnan(e_x))
    e_x = (e_x.astype(np.float64) * d1).astype(dt)                           # e_x_cov_x *= gauss_d1_
    cg = np.einsum("mij,mjk->mik", ci, PG)                                   # c_inv * point_gradient
    xcg = np.einsum("mi,mik->mk", xt, cg)                                    # (m, 6)
    g_pair = (e_x[:, None] * xcg).astype(np.float64)
    gg = np.einsum("mri,mrj->mij", PG, cg)                                   # point_gradient^T * c_inv * point_gradient
    xh_ij = np.einsum("mr,mirj->mij", xc, PH)                                # x_trans^T c_inv * point_hessian block i
    # hessian(i, j) += e * (-d2 * xcg(i) * xcg(j) + xh(i)(j) + gg(j, i))
    h_pair = (e_x[:, None, None] * (-gd2 * xcg[:, :, None] * xcg[:, None, :] + xh_ij
                                    + np.transpose(gg, (0, 2, 1)))).astype(np.float64)
    v = valid.astype(np.float64)
    score = float((score_inc.astype(np.float64) * v).sum())
    g = (g_pair * v[:, None]).sum(axis=0)
    H = (h_pair * v[:, None, None]).sum(axis=0)
    return score, g, H, m / float(source_f32.shape[0])


def _ndt_pose_matrix_f32(x):
    """Translation<float>(x0,x1,x2) * AngleAxis(x3, X) * AngleAxis(x4, Y) * AngleAxis(x5, Z), all float."""
    t = [F32(v) for v in x]
    cx, sx = np.cos(t[3]), np.sin(t[3]); cy, sy = np.cos(t[4]), np.sin(t[4]); cz, sz = np.cos(t[5]), np.sin(t[5])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=F32)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=F32)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=F32)
    T = np.eye(4, dtype=F32)
    T[:3, :3] = (rx @ ry) @ rz
    T[:3, 3] = t[:3]
    return T


def _transform_cloud_f32(T, pts):
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    return np.stack([((T[i, 0] * x + T[i, 1] * y) + T[i, 2] * z) + T[i, 3] for i in range(3)], axis=1).astype(F32)


def _euler_angles_012_f32(R):
    """Eigen 3.3 MatrixBase::eulerAngles(0, 1, 2) (Eigen/src/Geometry/EulerAngles.h), float."""
    r0 = math.atan2(float(R[1, 2]), float(R[2, 2]))
    c2 = math.hypot(float(R[0, 0]), float(R[0, 1]))
    if r0 > 0.0:
        r0 -= math.pi
        r1 = math.atan2(-float(R[0, 2]), -c2)
    else:
        r1 = math.atan2(-float(R[0, 2]), c2)
    s1, c1 = math.sin(r0), math.cos(r0)
    r2 = math.atan2(s1 * float(R[2, 0]) - c1 * float(R[1, 0]), c1 * float(R[1, 1]) - s1 * float(R[2, 1]))
    return [F32(-r0), F32(-r1), F32(-r2)]


def ndt_align(source_f32, target_f32, guess=None, resolution=1.0, step_size=0.1, outlier_ratio=0.55,
              transformation_epsilon=0.1, max_iterations=35, f64=False):
    """Ndt::Align: pclomp computeTransformation + pcl::Registration::getFitnessScore.  f64 = True with
    transformation_epsilon = 0.01: the stock PCL NDT stage of NdtWithGicp::Align (ndt_gicp.cc:44-47,82-88)."""
    from scipy.spatial import cKDTree
    src = np.ascontiguousarray(source_f32, dtype=F32)
    tgt = np.ascontiguousarray(target_f32, dtype=F32)
    grid = NdtGrid(tgt, resolution)
    final = np.eye(4, dtype=F32)
    output = src.copy()
    g4 = np.eye(4, dtype=F32) if guess is None else np.asarray(guess, dtype=np.float64).astype(F32)
    if not np.array_equal(g4, np.eye(4, dtype=F32)):
        final = g4
        output = _transform_cloud_f32(g4, output)
    rot = _euler_angles_012_f32(final[:3, :3])
    p = np.array([final[0, 3], final[1, 3], final[2, 3], rot[0], rot[1], rot[2]], dtype=np.float64)
    score, grad, hess, nbar = ndt_derivatives(grid, src, output, p, outlier_ratio, f64)
    nr_iterations, converged, evaluations, nb_sum = 0, False, 1, nbar
    while not converged:
        delta_p = np.linalg.lstsq(hess, -grad, rcond=None)[0]        # JacobiSVD(hessian).solve(-score_gradient)
        delta_p_norm = float(np.linalg.norm(delta_p))
        if delta_p_norm == 0 or delta_p_norm != delta_p_norm:
            break
        delta_p = delta_p / delta_p_norm
        # computeStepLengthMT with the loop that never runs
        d_phi_0 = -float(grad @ delta_p)
        if d_phi_0 >= 0:
            if d_phi_0 == 0:
                a_t = 0.0
                delta_p = delta_p * a_t
                break_now = True
            else:
                delta_p = -delta_p
                break_now = False
        else:
            break_now = False
        if not break_now:
            a_t = max(min(delta_p_norm, step_size), transformation_epsilon / 2)
            x_t = p + delta_p * a_t
            final = _ndt_pose_matrix_f32(x_t)
            output = _transform_cloud_f32(final, src)
            score, grad, hess, nbar = ndt_derivatives(grid, src, output, x_t, outlier_ratio, f64)
            evaluations += 1
            nb_sum += nbar
            delta_p = delta_p * a_t
        p = p + delta_p
        if nr_iterations > max_iterations or (nr_iterations and abs(a_t) < transformation_epsilon):
            converged = True
        nr_iterations += 1
    trans_probability = score / float(src.shape[0])
    # getFitnessScore(): mean squared distance of final * source to its exact nearest neighbour in the target
    moved = _transform_cloud_f32(final, src)
    dd, _ = cKDTree(tgt.astype(np.float64)).query(moved.astype(np.float64))
    fitness = float(np.mean(dd * dd))
    return {"result": final.astype(np.float64), "iterations": nr_iterations, "evaluations": evaluations,
            "fitness": fitness, "trans_probability": trans_probability,
            "mean_neighbors": nb_sum / evaluations}          # a diagnostic of this repo (mean over the evaluations)


# ------------------------------------------------------------------------------------ GICP pieces
# registrators/pclomp/gicp_omp_impl.hpp (the in-tree statement of the stock PCL GICP that NdtWithGicp runs):
# :59-131 computeCovariances, :419-463 correspondences + Mahalanobis matrices, :255-377 cost / gradient,
# :133-183 computeRDerivative, :516-527 applyState.
def gicp_covariances(cloud_f32, k=20, eps=1e-3):
    from scipy.spatial import cKDTree
    c = np.ascontiguousarray(cloud_f32, dtype=F32)
    _, nn = cKDTree(c.astype(np.float64)).query(c.astype(np.float64), k=k)
    out = np.zeros((c.shape[0], 3, 3))
    for i in range(c.shape[0]):
        pts = c[nn[i]]
        mean = pts.astype(np.float64).sum(axis=0) / k
        prod = (pts[:, :, None] * pts[:, None, :]).astype(np.float64)       # float products, double sums
        cov = prod.sum(axis=0) / k - np.outer(mean, mean)
        w, V = np.linalg.eigh(cov)                                           # ascending: column 0 = smallest
        out[i] = np.outer(V[:, 2], V[:, 2]) + np.outer(V[:, 1], V[:, 1]) + eps * np.outer(V[:, 0], V[:, 0])
    return out


def _rot_zyx_f32(x3, x4, x5):
    cx, sx = np.cos(F32(x3)), np.sin(F32(x3)); cy, sy = np.cos(F32(x4)), np.sin(F32(x4))
    cz, sz = np.cos(F32(x5)), np.sin(F32(x5))
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=F32)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=F32)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=F32)
    return (rz @ ry) @ rx


def gicp_apply_state(T_f32, x):
    """applyState: t.R = Rz(x5) Ry(x4) Rx(x3) * t.R ; t.t += x[0:3]   (all float)."""
    T = np.array(T_f32, dtype=F32)
    T[:3, :3] = _rot_zyx_f32(x[3], x[4], x[5]) @ T[:3, :3]
    T[:3, 3] += np.asarray(x[:3]).astype(F32)
    return T


def gicp_correspond(src_f32, tgt_f32, guess_f32, transformation_f32, cov_s, cov_t, corr_dist=5.0):
    """-> (source indices, target indices, M_i for every source point (identity without a correspondence))."""
    from scipy.spatial import cKDTree
    src = np.ascontiguousarray(src_f32, dtype=F32); tgt = np.ascontiguousarray(tgt_f32, dtype=F32)
    G = np.asarray(guess_f32, dtype=F32); Tm = np.asarray(transformation_f32, dtype=F32)
    R = (Tm.astype(np.float64) @ G.astype(np.float64))[:3, :3]             # transform_R formed in double
    query = _transform_cloud_f32(Tm, _transform_cloud_f32(G, src))
    _, nn = cKDTree(tgt.astype(np.float64)).query(query.astype(np.float64))
    d = query - tgt[nn]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]        # float squared distance
    ok = d2.astype(np.float64) < corr_dist * corr_dist
    M = np.tile(np.eye(3), (src.shape[0], 1, 1))
    si = np.flatnonzero(ok)
    ti = nn[si]
    M[si] = np.linalg.inv(R @ cov_s[si] @ R.T + cov_t[ti])
    return si, ti, M


def gicp_cost(src_f32, tgt_f32, base_f32, M, si, ti, x):
    """(f, g[6]) of OptimizationFunctorWithIndices at state x."""
    src = np.ascontiguousarray(src_f32, dtype=F32); tgt = np.ascontiguousarray(tgt_f32, dtype=F32)
    base = np.asarray(base_f32, dtype=F32)
    T = gicp_apply_state(base, x)
    pp = _transform_cloud_f32(T, src[si])
    res = (pp - tgt[ti]).astype(np.float64)                                  # float subtraction, then double
    temp = np.einsum("mij,mj->mi", M[si], res)
    m = si.size
    f = float(np.einsum("mi,mi->", res, temp)) / m
    g = np.zeros(6)
    g[:3] = temp.sum(axis=0) * (2.0 / m)
    pb = _transform_cloud_f32(base, src[si]).astype(np.float64)
    Rm = (pb.T @ temp) * (2.0 / m)
    phi, theta, psi = x[3], x[4], x[5]
    cphi, sphi, cth, sth, cpsi, spsi = math.cos(phi), math.sin(phi), math.cos(theta), math.sin(theta), math.cos(psi), math.sin(psi)
    # d(Rz(psi) Ry(theta) Rx(phi)) / d angle, written out from the rotation itself (not from the table in the source)
    def rot(ph, th, ps):
        cx, sx, cy, sy, cz, sz = math.cos(ph), math.sin(ph), math.cos(th), math.sin(th), math.cos(ps), math.sin(ps)
        return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]]) @ np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
                @ np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    h = 1e-6
    for k, dv in enumerate(np.eye(3)):
        dR = (rot(phi + h * dv[0], theta + h * dv[1], psi + h * dv[2])
              - rot(phi - h * dv[0], theta - h * dv[1], psi - h * dv[2])) / (2 * h)
        # matricesInnerProd(dR, Rm) = sum_ij dR(j,i) Rm(i,j) = tr(dR Rm) = sum_k temp_k^T dR (base p_k): df/dangle
        g[3 + k] = float(np.sum(dR * Rm.T))
    return f, g


# ------------------------------------------------------------------------------------ IcpUsingPointMatcher
# The reference's type-1 matcher is libpointmatcher 1.3.1 (external, `PointMatcher<float>`) driven by the default
# chain of registrators/icp_pointmatcher.cc:166-247; its Align() and score pass are :95-148.  libpointmatcher is
# not in the reference tree; the modules are restated from the published sources of that release:
#   RandomSamplingDataPointsFilter (prob 0.9):   keep point i iff (float)rand() / (float)RAND_MAX < prob
#   SamplingSurfaceNormalDataPointsFilter (knn 7, samplingMethod 1): the recursion the reference author ported as
#       CalculateNormals (cloud_types.cc:73-144), but the normal is the eigenvector of the smallest eigenvalue of
#       the box's covariance, and the kept point is the box mean
#   KDTreeMatcher (knn 1, epsilon 3.16), TrimmedDistOutlierFilter (0.7), PointToPlaneErrorMinimizer,
#   CounterTransformationChecker (150), DifferentialTransformationChecker (0.001 rad, 0.01 m, smoothLength 4):
#       the chain icp_fast.cc restates in double.
# `rand()` is glibc's TYPE_3 additive feedback generator (r[i] = r[i-3] + r[i-31], 310 outputs discarded,
# result >> 1), default seed 1 — a process-global stream, so only the FIRST Align of a process is reproducible.
class GlibcRand:
    RAND_MAX = 2147483647

    def __init__(self, seed=1):
        r = [0] * 34
        r[0] = seed if seed != 0 else 1
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)          # 16807 * r mod (2^31 - 1) without overflow, as glibc does
            word = 16807 * lo - 2836 * hi
            if word < 0:
                word += 2147483647
            r[i] = word
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = r
        for _ in range(310):
            self._next()

    def _next(self):
        v = (self.r[-31] + self.r[-3]) & 0xFFFFFFFF
        self.r.append(v)
        if len(self.r) > 64:
            del self.r[:len(self.r) - 34]
        return v

    def rand(self):
        return self._next() >> 1


def pm_random_sampling_mask(n, prob=0.9, rng=None):
    rng = rng or GlibcRand(1)
    p = F32(prob)
    denom = F32(GlibcRand.RAND_MAX)                      # (float)RAND_MAX = 2147483648.0f
    return np.array([F32(rng.rand()) / denom < p for _ in range(n)], dtype=bool)


def pm_sampling_surface_normal(points_f32, knn=7):
    """-> (means (M,3) f32, normals (M,3) f32) in ascending order of the box's smallest member index."""
    pts = np.ascontiguousarray(points_f32, dtype=F32)
    rows = [tuple(float(x) for x in p) for p in pts]
    kept = []

    def fuse(idx):
        d = pts[sorted(idx)].T                                       # 3 x count, float
        mean = d.sum(axis=1, dtype=F32) / F32(d.shape[1])
        nn = d - mean[:, None]
        c = (nn @ nn.T).astype(F32)
        if np.linalg.matrix_rank(c.astype(np.float64), tol=float(np.finfo(F32).eps) * 3 * np.abs(c).max()) + 1 < 3:
            return
        w, v = np.linalg.eigh(c)                                      # EigenSolver on a symmetric matrix: real spectrum
        kept.append((min(idx), mean, v[:, int(np.argmin(w))]))

    def build(idx, min_values, max_values):
        if len(idx) <= knn:
            fuse(idx)
            return
        cut_dim = _arg_max([max_values[r] - min_values[r] for r in range(3)])
        left, right = _median_split(rows, idx, cut_dim)
        cut_val = rows[right[0]][cut_dim]
        left_max = list(max_values); left_max[cut_dim] = cut_val
        right_min = list(min_values); right_min[cut_dim] = cut_val
        build(left, min_values, left_max)
        build(right, right_min, max_values)

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    build(list(range(len(rows))), [float(x) for x in pts.min(axis=0)], [float(x) for x in pts.max(axis=0)])
    kept.sort(key=lambda t: t[0])
    return (np.array([k[1] for k in kept], dtype=F32).reshape(-1, 3),
            np.array([k[2] for k in kept], dtype=F32).reshape(-1, 3))


def icp_pm_literal(source_f32, target_f32, guess=None, rng=None, knn_full=None):
    """IcpUsingPointMatcher::Align with the literal data filters; the iteration itself in double (the float
    arithmetic of PointMatcher<float> is NOT reproduced: its effect is at the 1e-5 m level, the filters' at 1e-3 m).
    knn_full(target, query) -> (ids, d2) is used for the score pass over the unfiltered clouds (defaults to PyNabo)."""
    src = np.ascontiguousarray(source_f32, dtype=F32)
    tgt = np.ascontiguousarray(target_f32, dtype=F32)
    keep = pm_random_sampling_mask(src.shape[0], 0.9, rng)
    tp, tn = pm_sampling_surface_normal(tgt, 7)
    out = icp_fast_align(src[keep].astype(np.float64), tp.astype(np.float64), tn.astype(np.float64), guess,
                         max_iteration=150)
    out["n_source"], out["n_target"] = int(keep.sum()), int(tp.shape[0])
    # score pass (icp_pointmatcher.cc:101-148): unfiltered reading moved by the result, matched against the
    # unfiltered reference, trimmed at 0.7, exp(-mean distance); Align() is false below 0.6
    moved = _apply_transform(out["result"], src.astype(np.float64))
    t64 = tgt.astype(np.float64)
    ids, d2 = (knn_full(t64, moved) if knn_full else PyNabo(t64).knn1(moved, 3.16))
    limit = _quantile_limit(d2, 0.7)
    kept = d2[(d2 != INF) & (d2 <= limit)]
    out["icp_score"] = out["score"]
    out["score"] = math.exp(-float(np.sqrt(kept).sum()) / kept.size)
    out["ok"] = out["score"] >= 0.6
    return out
