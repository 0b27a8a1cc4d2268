// ORACLE — TEST INFRASTRUCTURE ONLY (see sm_oracle.h).
//
// descriptor::M2dp (descriptor/m2dp.cc:37-172, descriptor/m2dp.h:45-84), the loop-closure
// descriptor computed from every finished submap.  Restated with the reference's own types and
// operation order; its peculiarities are kept on purpose:
//   * preProcess (:45-70): pcl::PCA<pcl::PointXYZ> in single precision — centroid accumulated in
//     float in input order, covariance of the demeaned cloud / (n - 1), eigenvectors by descending
//     eigenvalue, third axis = first x second; project() = eigenvectors^T (x - mean); points with
//     ||.|| > max_distance are dropped; l = ceil(sqrt(max_distance / r));
//   * singleViewProcess (:72-127): m = (cos t cos p, cos t sin p, sin t) in float;
//     x_axis = e_x - |e_x . m| m,  y_axis = m x x_axis  (not normalised);  the projected point is
//     (|p . x_axis|, |p . y_axis|) — `(p.transpose() * axis).norm()` of a 1x1 product is an
//     ABSOLUTE VALUE, so only the first quadrant of the polar grid is ever filled;
//     length = float norm, angle = atan2 of two floats (float overload), bins in double;
//   * setInputCloud (:129-153): A (p*q x l*t, float counts), JacobiSVD, descriptor = [u1; v1].
// PARITY UNPINNED (the reference has no M2dp test and cannot be built here) and, beyond that, two
// SIGN conventions live inside Eigen (SelfAdjointEigenSolver's eigenvector signs, JacobiSVD's
// singular-vector signs) and cannot be pinned from this tree.  Conventions used here and by the
// CUDA path: each of the first two PCA axes is oriented so that its component of largest magnitude
// is positive; (u1, v1) is oriented so that the component of u1 of largest magnitude is positive.
// matchTwoM2dpDescriptors (:155-170, |Pearson correlation|) does not depend on the second one.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "sm_oracle.h"

namespace {

// symmetric eigen decomposition by cyclic Jacobi (double), eigenvalues descending, columns of V
void jacobi_eig(int n, std::vector<double>& A, std::vector<double>& V, std::vector<double>& w) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (fabs(apq) < 1e-300) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq; A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk; A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] > A[(size_t)b * n + b]; });
  std::vector<double> V2((size_t)n * n);
  w.resize(n);
  for (int j = 0; j < n; ++j) {
    w[j] = A[(size_t)order[j] * n + order[j]];
    for (int k = 0; k < n; ++k) V2[(size_t)k * n + j] = V[(size_t)k * n + order[j]];
  }
  V.swap(V2);
}

}  // namespace

extern "C" {

int64_t sm_oracle_m2dp_dims(double r, double max_distance, int32_t t, int32_t p, int32_t q, int32_t* l_out) {
  if (r < 1.e-6 || t <= 0 || p <= 0 || q <= 0) return -1;
  const int32_t l = (int32_t)ceil(sqrt(max_distance / r));
  if (l_out) *l_out = l;
  return (int64_t)p * q + (int64_t)l * t;
}

// points: packed float xyz.  descriptor: p*q + l*t floats.  A_out (optional): p*q x l*t counts,
// row-major.  axes_out (optional): 12 floats = mean (3) + the three PCA axes (columns, 9).
// Returns the descriptor length, 0 for an empty cloud (setInputCloud returns false), -1 for r < 1e-6.
int64_t sm_oracle_m2dp(const float* points, int64_t n, double r, double max_distance, int32_t t_, int32_t p_,
                       int32_t q_, float* descriptor, int32_t* A_out, float* axes_out) {
  if (n <= 0) return 0;                                   // m2dp.cc:130-133
  int32_t l_ = 0;
  const int64_t len = sm_oracle_m2dp_dims(r, max_distance, t_, p_, q_, &l_);
  if (len < 0) return -1;                                 // m2dp.cc:64-67 (descriptor left unset)
  // ---- pcl::PCA (float): compute3DCentroid, demeanPointCloud, covariance / (n - 1) ----------
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (int64_t i = 0; i < n; ++i) { cx += points[3 * i]; cy += points[3 * i + 1]; cz += points[3 * i + 2]; }
  cx /= (float)n; cy /= (float)n; cz /= (float)n;
  float c[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    const float x = points[3 * i] - cx, y = points[3 * i + 1] - cy, z = points[3 * i + 2] - cz;
    c[0] += x * x; c[1] += x * y; c[2] += x * z; c[3] += y * y; c[4] += y * z; c[5] += z * z;
  }
  const float denom = (float)(n - 1);
  std::vector<double> C = {c[0] / denom, c[1] / denom, c[2] / denom, c[1] / denom, c[3] / denom, c[4] / denom,
                           c[2] / denom, c[4] / denom, c[5] / denom};
  std::vector<double> V, w;
  jacobi_eig(3, C, V, w);
  float E[9];                                             // column j = axis j
  for (int j = 0; j < 2; ++j) {
    int big = 0;
    for (int k = 1; k < 3; ++k) if (fabs(V[(size_t)k * 3 + j]) > fabs(V[(size_t)big * 3 + j])) big = k;
    const double sgn = V[(size_t)big * 3 + j] < 0 ? -1.0 : 1.0;
    for (int k = 0; k < 3; ++k) E[k * 3 + j] = (float)(sgn * V[(size_t)k * 3 + j]);
  }
  E[0 * 3 + 2] = E[1 * 3 + 0] * E[2 * 3 + 1] - E[2 * 3 + 0] * E[1 * 3 + 1];   // col2 = col0 x col1
  E[1 * 3 + 2] = E[2 * 3 + 0] * E[0 * 3 + 1] - E[0 * 3 + 0] * E[2 * 3 + 1];
  E[2 * 3 + 2] = E[0 * 3 + 0] * E[1 * 3 + 1] - E[1 * 3 + 0] * E[0 * 3 + 1];
  if (axes_out) { axes_out[0] = cx; axes_out[1] = cy; axes_out[2] = cz; memcpy(axes_out + 3, E, sizeof(E)); }
  // ---- project + distance gate (m2dp.cc:55-62) ------------------------------------------------
  std::vector<float> P;
  P.reserve((size_t)n * 3);
  for (int64_t i = 0; i < n; ++i) {
    const float x = points[3 * i] - cx, y = points[3 * i + 1] - cy, z = points[3 * i + 2] - cz;
    const float px = E[0] * x + E[3] * y + E[6] * z, py = E[1] * x + E[4] * y + E[7] * z,
                pz = E[2] * x + E[5] * y + E[8] * z;
    const double d = (double)sqrtf(px * px + py * py + pz * pz);    // getLength<PointXYZ>: all float
    if (d <= max_distance) { P.push_back(px); P.push_back(py); P.push_back(pz); }
  }
  const int64_t m = (int64_t)P.size() / 3;
  // ---- views (m2dp.cc:72-127, 139-146) ----------------------------------------------------------
  const int rows = p_ * q_, cols = l_ * t_;
  std::vector<double> A((size_t)rows * cols, 0.0);
  const double theta_step = M_PI / p_, phi_step = M_PI_2 / q_, two_pi = M_PI * 2., angle_step = two_pi / t_;
  for (int pi = 0; pi < p_; ++pi)
    for (int qi = 0; qi < q_; ++qi) {
      const double theta = pi * theta_step, phi = qi * phi_step;
      const float mx = (float)(cos(theta) * cos(phi)), my = (float)(cos(theta) * sin(phi)), mz = (float)sin(theta);
      const float s = fabsf(mx);                               // (e_x^T m).norm()
      const float xa[3] = {1.f - s * mx, 0.f - s * my, 0.f - s * mz};
      const float ya[3] = {my * xa[2] - mz * xa[1], mz * xa[0] - mx * xa[2], mx * xa[1] - my * xa[0]};
      double* row = &A[(size_t)(pi * q_ + qi) * cols];
      for (int64_t i = 0; i < m; ++i) {
        const float x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
        const float u = fabsf(x * xa[0] + y * xa[1] + z * xa[2]), v = fabsf(x * ya[0] + y * ya[1] + z * ya[2]);
        const double length = (double)sqrtf(u * u + v * v);    // Vector2f::norm()
        double angle = (double)atan2f(v, u);
        if (angle < 0.) angle += two_pi;
        int32_t li = (int32_t)floor(sqrt(length / r));
        if (li > l_ - 1) li = l_ - 1;
        int32_t ti = (int32_t)floor(angle / angle_step);
        if (ti > t_ - 1) ti = t_ - 1;
        row[li * t_ + ti] += 1.0;
      }
    }
  if (A_out) for (size_t k = 0; k < A.size(); ++k) A_out[k] = (int32_t)A[k];
  // ---- first singular vectors: u1 = top eigenvector of A A^T, v1 = A^T u1 / sigma1 ------------------
  std::vector<double> G((size_t)rows * rows, 0.0);
  for (int a = 0; a < rows; ++a)
    for (int b = a; b < rows; ++b) {
      double s = 0.0;
      for (int k = 0; k < cols; ++k) s += A[(size_t)a * cols + k] * A[(size_t)b * cols + k];
      G[(size_t)a * rows + b] = s; G[(size_t)b * rows + a] = s;
    }
  std::vector<double> U, ev;
  jacobi_eig(rows, G, U, ev);
  int big = 0;
  for (int k = 1; k < rows; ++k) if (fabs(U[(size_t)k * rows]) > fabs(U[(size_t)big * rows])) big = k;
  const double sgn = U[(size_t)big * rows] < 0 ? -1.0 : 1.0;
  const double sigma = sqrt(std::max(ev[0], 0.0));
  for (int k = 0; k < rows; ++k) descriptor[k] = (float)(sgn * U[(size_t)k * rows]);
  for (int k = 0; k < cols; ++k) {
    double s = 0.0;
    for (int a = 0; a < rows; ++a) s += A[(size_t)a * cols + k] * (sgn * U[(size_t)a * rows]);
    descriptor[rows + k] = sigma > 0 ? (float)(s / sigma) : 0.f;
  }
  return len;
}

// matchTwoM2dpDescriptors (m2dp.cc:155-170); -1 where the reference prints an error
double sm_oracle_m2dp_match(const float* P, const float* Q, int64_t n) {
  if (n < 10) return -1.;
  float pq = 0.f, pp = 0.f, qq = 0.f, sp = 0.f, sq = 0.f;   // Eigen VectorXf dot / sum: float
  for (int64_t i = 0; i < n; ++i) { pq += P[i] * Q[i]; pp += P[i] * P[i]; qq += Q[i] * Q[i]; sp += P[i]; sq += Q[i]; }
  const double N = (double)n;
  const double score = (N * pq - (double)(sp * sq)) / sqrt((N * pp - pow((double)sp, 2)) * (N * qq - pow((double)sq, 2)));
  return fabs(score);
}

}  // extern "C"
