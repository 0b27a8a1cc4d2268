// Device-side 6x6 / 3x3 dense algebra used by the ICP finish step and by target prep.
// Follows the Eigen 3.3 routines the reference calls (FullPivHouseholderQR::rank /
// isInvertible, LLT::solve, JacobiSVD::solve on the symmetric normal matrix,
// PartialPivLU::inverse, AngleAxis::toRotationMatrix, Quaternion(Matrix3),
// angularDistance) — call sites: registrators/icp_fast.cc:204-254,307-321,377-405 and
// builder/data/cloud_types.cc:89-93.  Row-major m[r*N+c] unless noted.
#ifndef SM_B200_LINALG_DEV_CUH_
#define SM_B200_LINALG_DEV_CUH_

#include <float.h>

namespace smb {
namespace la {

constexpr double kEps = 2.220446049250313e-16;

template <int N>
struct PivQR {
  double R[N * N];
  double Qt[N * N];
  int perm[N];
  int nonzero_pivots;
  double maxpivot;

  __host__ __device__ void compute(const double* A) {
    const double precision = kEps * (double)N;
    for (int i = 0; i < N * N; ++i) { R[i] = A[i]; Qt[i] = 0.0; }
    for (int i = 0; i < N; ++i) { Qt[i * N + i] = 1.0; perm[i] = i; }
    nonzero_pivots = N;
    maxpivot = 0.0;
    double biggest = 0.0;
    for (int k = 0; k < N; ++k) {
      int rb = k, cb = k;
      double best = -1.0;
      for (int c = k; c < N; ++c)        // column-major visit, strict '>' (Eigen maxCoeff)
        for (int r = k; r < N; ++r) {
          const double v = fabs(R[r * N + c]);
          if (v > best) { best = v; rb = r; cb = c; }
        }
      if (k == 0) biggest = best;
      if (fabs(best) <= fabs(biggest) * precision) { nonzero_pivots = k; break; }
      if (rb != k)
        for (int c = 0; c < N; ++c) {
          double t = R[k * N + c]; R[k * N + c] = R[rb * N + c]; R[rb * N + c] = t;
          t = Qt[k * N + c]; Qt[k * N + c] = Qt[rb * N + c]; Qt[rb * N + c] = t;
        }
      if (cb != k) {
        for (int r = 0; r < N; ++r) {
          const double t = R[r * N + k]; R[r * N + k] = R[r * N + cb]; R[r * N + cb] = t;
        }
        const int t = perm[k]; perm[k] = perm[cb]; perm[cb] = t;
      }
      double tail_sq = 0.0;
      for (int r = k + 1; r < N; ++r) tail_sq += R[r * N + k] * R[r * N + k];
      const double c0 = R[k * N + k];
      double beta, tau, v[N];
      for (int r = 0; r < N; ++r) v[r] = 0.0;
      if (tail_sq <= DBL_MIN) { tau = 0.0; beta = c0; }
      else {
        beta = sqrt(c0 * c0 + tail_sq);
        if (c0 >= 0.0) beta = -beta;
        for (int r = k + 1; r < N; ++r) v[r] = R[r * N + k] / (c0 - beta);
        tau = (beta - c0) / beta;
      }
      v[k] = 1.0;
      if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
      R[k * N + k] = beta;
      for (int r = k + 1; r < N; ++r) R[r * N + k] = 0.0;
      if (tau != 0.0) {
        for (int c = k + 1; c < N; ++c) {
          double dot = 0.0;
          for (int r = k; r < N; ++r) dot += v[r] * R[r * N + c];
          for (int r = k; r < N; ++r) R[r * N + c] -= tau * v[r] * dot;
        }
        for (int c = 0; c < N; ++c) {
          double dot = 0.0;
          for (int r = k; r < N; ++r) dot += v[r] * Qt[r * N + c];
          for (int r = k; r < N; ++r) Qt[r * N + c] -= tau * v[r] * dot;
        }
      }
    }
  }
  __host__ __device__ int rank() const {
    const double thr = fabs(maxpivot) * (kEps * (double)N);
    int rk = 0;
    for (int i = 0; i < nonzero_pivots; ++i) rk += (fabs(R[i * N + i]) > thr) ? 1 : 0;
    return rk;
  }
};

// Cholesky solve, n <= 6
__host__ __device__ inline void llt_solve(const double* A, const double* b, double* x, int n) {
  double L[36];
  for (int i = 0; i < n * n; ++i) L[i] = 0.0;
  for (int k = 0; k < n; ++k) {
    double d = A[k * n + k];
    for (int j = 0; j < k; ++j) d -= L[k * n + j] * L[k * n + j];
    const double lkk = sqrt(d);
    L[k * n + k] = lkk;
    for (int i = k + 1; i < n; ++i) {
      double s = A[i * n + k];
      for (int j = 0; j < k; ++j) s -= L[i * n + j] * L[k * n + j];
      L[i * n + k] = s / lkk;
    }
  }
  double y[6];
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int j = 0; j < i; ++j) s -= L[i * n + j] * y[j];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = y[i];
    for (int j = i + 1; j < n; ++j) s -= L[j * n + i] * x[j];
    x[i] = s / L[i * n + i];
  }
}

// symmetric Jacobi eigen-decomposition, n <= 6 (V columns = eigenvectors)
__host__ __device__ inline void jacobi_eig_sym(const double* Ain, int n, double* w, double* V) {
  double A[36];
  for (int i = 0; i < n * n; ++i) { A[i] = Ain[i]; V[i] = 0.0; }
  for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
    if (off == 0.0) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

__host__ __device__ inline void sym_svd_solve(const double* A, const double* b, double* x, int n) {
  double w[6], V[36];
  jacobi_eig_sym(A, n, w, V);
  double svmax = 0.0;
  for (int i = 0; i < n; ++i) svmax = fmax(svmax, fabs(w[i]));
  const double thr = fmax(svmax * (double)n * kEps, DBL_MIN);
  for (int i = 0; i < n; ++i) x[i] = 0.0;
  for (int k = 0; k < n; ++k) {
    if (!(fabs(w[k]) > thr)) continue;
    double dot = 0.0;
    for (int i = 0; i < n; ++i) dot += V[i * n + k] * b[i];
    dot /= w[k];
    for (int i = 0; i < n; ++i) x[i] += V[i * n + k] * dot;
  }
}

// icp_fast.cc:204-254.  path: 0 LLT, 1 rank-reduced min-norm, 2 SVD fallback.
__host__ __device__ inline int solve_possibly_underdetermined(const double* A, const double* b, double* x) {
  PivQR<6> qr;
  qr.compute(A);
  const int rank = qr.rank();
  if (rank == 6) { llt_solve(A, b, x, 6); return 0; }
  double QA[36], R1[36], Qb[6], G[36], y[6], xp[6];
  for (int r = 0; r < rank; ++r)
    for (int k = 0; k < 6; ++k) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += qr.Qt[r * 6 + m] * A[m * 6 + k];
      QA[r * 6 + k] = s;
    }
  for (int r = 0; r < rank; ++r) {
    for (int c = 0; c < 6; ++c) R1[r * 6 + c] = QA[r * 6 + qr.perm[c]];
    double s = 0.0;
    for (int m = 0; m < 6; ++m) s += qr.Qt[r * 6 + m] * b[m];
    Qb[r] = s;
  }
  for (int r = 0; r < rank; ++r)
    for (int c = 0; c < rank; ++c) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s += R1[r * 6 + k] * R1[c * 6 + k];
      G[r * rank + c] = s;
    }
  for (int i = 0; i < 6; ++i) { y[i] = 0.0; xp[i] = 0.0; x[i] = 0.0; }
  if (rank > 0) llt_solve(G, Qb, y, rank);
  for (int c = 0; c < 6; ++c)
    for (int r = 0; r < rank; ++r)
      if (c >= r) xp[c] += R1[r * 6 + c] * y[r];
  for (int j = 0; j < 6; ++j) x[qr.perm[j]] = xp[j];
  double nb = 0.0, nax = 0.0, nd = 0.0;
  for (int r = 0; r < 6; ++r) {
    double ax = 0.0;
    for (int c = 0; c < 6; ++c) ax += A[r * 6 + c] * x[c];
    nb += b[r] * b[r]; nax += ax * ax; nd += (b[r] - ax) * (b[r] - ax);
  }
  if (!(nd <= 1e-10 * fmin(nb, nax))) { sym_svd_solve(A, b, x, 6); return 2; }
  return 1;
}

// 3x3 inverse via partial-pivot LU (dynamic MatrixXd::inverse path, cloud_types.cc:93)
__host__ __device__ inline void lu_inverse3(const double* M, double* inv) {
  double lu[9];
  int piv[3] = {0, 1, 2};
  for (int i = 0; i < 9; ++i) lu[i] = M[i];
  for (int k = 0; k < 3; ++k) {
    int rb = k;
    double best = fabs(lu[k * 3 + k]);
    for (int r = k + 1; r < 3; ++r)
      if (fabs(lu[r * 3 + k]) > best) { best = fabs(lu[r * 3 + k]); rb = r; }
    if (rb != k) {
      for (int c = 0; c < 3; ++c) { const double t = lu[k * 3 + c]; lu[k * 3 + c] = lu[rb * 3 + c]; lu[rb * 3 + c] = t; }
      const int t = piv[k]; piv[k] = piv[rb]; piv[rb] = t;
    }
    for (int r = k + 1; r < 3; ++r) {
      lu[r * 3 + k] /= lu[k * 3 + k];
      for (int c = k + 1; c < 3; ++c) lu[r * 3 + c] -= lu[r * 3 + k] * lu[k * 3 + c];
    }
  }
  for (int col = 0; col < 3; ++col) {
    double y[3], x[3];
    for (int r = 0; r < 3; ++r) {
      double s = (piv[r] == col) ? 1.0 : 0.0;
      for (int j = 0; j < r; ++j) s -= lu[r * 3 + j] * y[j];
      y[r] = s;
    }
    for (int r = 2; r >= 0; --r) {
      double s = y[r];
      for (int j = r + 1; j < 3; ++j) s -= lu[r * 3 + j] * x[j];
      x[r] = s / lu[r * 3 + r];
    }
    for (int r = 0; r < 3; ++r) inv[r * 3 + col] = x[r];
  }
}

__host__ __device__ inline void angle_axis_to_rotation(double angle, const double* axis, double* R) {
  double s, c;
  sincos(angle, &s, &c);
  const double sa0 = s * axis[0], sa1 = s * axis[1], sa2 = s * axis[2];
  const double ca0 = (1.0 - c) * axis[0], ca1 = (1.0 - c) * axis[1], ca2 = (1.0 - c) * axis[2];
  double t;
  t = ca0 * axis[1]; R[1] = t - sa2; R[3] = t + sa2;
  t = ca0 * axis[2]; R[2] = t + sa1; R[6] = t - sa1;
  t = ca1 * axis[2]; R[5] = t - sa0; R[7] = t + sa0;
  R[0] = ca0 * axis[0] + c; R[4] = ca1 * axis[1] + c; R[8] = ca2 * axis[2] + c;
}

__host__ __device__ inline void rotation_to_quaternion(const double* m, double* q) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t; q[2] = (m[2] - m[6]) * t; q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    q[1 + i] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[k * 3 + j] - m[j * 3 + k]) * t;
    q[1 + j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
    q[1 + k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
  }
}

__host__ __device__ inline double quaternion_angular_distance(const double* a, const double* b) {
  const double bw = b[0], bx = -b[1], by = -b[2], bz = -b[3];
  const double w = a[0] * bw - a[1] * bx - a[2] * by - a[3] * bz;
  const double x = a[0] * bx + a[1] * bw + a[2] * bz - a[3] * by;
  const double y = a[0] * by + a[2] * bw + a[3] * bx - a[1] * bz;
  const double z = a[0] * bz + a[3] * bw + a[1] * by - a[2] * bx;
  return 2.0 * atan2(sqrt(x * x + y * y + z * z), fabs(w));
}

// column-major 4x4 product, k-order accumulation
__host__ __device__ inline void mul4(const double* A, const double* B, double* C) {
  double out[16];
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      double s = A[i] * B[j * 4];
      s += A[i + 4] * B[1 + j * 4];
      s += A[i + 8] * B[2 + j * 4];
      s += A[i + 12] * B[3 + j * 4];
      out[i + j * 4] = s;
    }
  for (int i = 0; i < 16; ++i) C[i] = out[i];
}

}  // namespace la
}  // namespace smb

#endif  // SM_B200_LINALG_DEV_CUH_
