// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// Scalar restatement of the per-point / per-leaf arithmetic of the vendored pclomp NDT, with
// every floating-point operation written out in a fixed order (the reference is Eigen
// expression code whose evaluation order depends on Eigen version and SIMD level; this file
// IS the order the oracle defines).  Compiled with -ffp-contract=off.
//   * leaf statistics      voxel_grid_covariance_omp_impl.hpp:209-263, 283-366
//   * angular derivatives  ndt_omp_impl.hpp:288-393
//   * point derivatives    ndt_omp_impl.hpp:397-438 (the float overload used by the OMP loop)
//   * updateDerivatives    ndt_omp_impl.hpp:483-535
#ifndef ORACLE_NDT_MATH_H_
#define ORACLE_NDT_MATH_H_

#include <cmath>
#include <limits>

#include "linalg.h"

namespace sm_oracle {

struct NdtGauss { double d1, d2, d3; };

// ndt_omp_impl.hpp:86-93 (eq. 6.8 [Magnusson 2009]); resolution is a float member.
inline void ndt_gauss_constants(double outlier_ratio, float resolution, NdtGauss* g) {
  const double c1 = 10.0 * (1 - outlier_ratio);
  const double c2 = outlier_ratio / std::pow((double)resolution, 3);
  g->d3 = -std::log(c2);
  g->d1 = -std::log(c1 + c2) - g->d3;
  g->d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - g->d3) / g->d1);
}

// voxel index of a point (_impl.hpp:218-223): float multiply, float floor, float subtract.
inline int ndt_voxel_index(float x, float y, float z, float inv_leaf, const int min_b[3],
                           const int mul[3]) {
  const int i0 = (int)(std::floor(x * inv_leaf) - (float)min_b[0]);
  const int i1 = (int)(std::floor(y * inv_leaf) - (float)min_b[1]);
  const int i2 = (int)(std::floor(z * inv_leaf) - (float)min_b[2]);
  return i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
}

// squared distance in single precision, x then y then z (FLANN L2_Simple)
inline float ndt_dist2f(float qx, float qy, float qz, const float* p) {
  const float dx = qx - p[0], dy = qy - p[1], dz = qz - p[2];
  return (dx * dx + dy * dy) + dz * dz;
}

inline void inverse3_cofactor(const double* m, double* inv) {   // Eigen fixed 3x3 inverse
  const double c00 = m[4] * m[8] - m[5] * m[7];
  const double c10 = m[5] * m[6] - m[3] * m[8];   // cofactor(1,0) = m12*m20 - m10*m22
  const double c20 = m[3] * m[7] - m[4] * m[6];
  const double det = (c00 * m[0] + c10 * m[1]) + c20 * m[2];
  const double invdet = 1.0 / det;
  inv[0] = c00 * invdet; inv[3] = c10 * invdet; inv[6] = c20 * invdet;
  inv[1] = (m[7] * m[2] - m[8] * m[1]) * invdet;
  inv[4] = (m[8] * m[0] - m[6] * m[2]) * invdet;
  inv[7] = (m[6] * m[1] - m[7] * m[0]) * invdet;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}

struct NdtLeafOut {
  int nr_points;
  int searchable;     // centroid is in the search cloud (n >= min_points)
  double mean[3];
  double icov[9];     // zero unless the leaf is valid
  float centroid[3];
};

// second pass of applyFilter for one leaf (_impl.hpp:283-366).  cov_acc is the x*x^T
// accumulator INCLUDING the identity it starts from; centroid_sum is the float sum.
inline void ndt_finalize_leaf(int n, const double* sum, const double* cov_acc,
                              const float* centroid_sum, int min_points, double eig_mult,
                              NdtLeafOut* out) {
  out->nr_points = n;
  out->searchable = 0;
  for (int k = 0; k < 9; ++k) out->icov[k] = 0.0;
  for (int d = 0; d < 3; ++d) {
    out->centroid[d] = centroid_sum[d] / (float)n;
    out->mean[d] = sum[d] / (double)n;
  }
  if (n < min_points) return;
  out->searchable = 1;
  double cov[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      cov[r * 3 + c] = (cov_acc[r * 3 + c] - 2 * (sum[r] * out->mean[c])) / (double)n + out->mean[r] * out->mean[c];
  const double scale = (n - 1.0) / n;
  for (int k = 0; k < 9; ++k) cov[k] *= scale;
  double w[3], V[9];
  JacobiEigenSym(cov, 3, w, V);
  int order[3] = {0, 1, 2};   // ascending eigenvalues (SelfAdjointEigenSolver convention)
  for (int a = 0; a < 3; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (w[order[b]] < w[order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
  double ev[3], E[9];
  for (int a = 0; a < 3; ++a) {
    ev[a] = w[order[a]];
    for (int r = 0; r < 3; ++r) E[r * 3 + a] = V[r * 3 + order[a]];
  }
  if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) { out->nr_points = -1; return; }
  const double min_ev = eig_mult * ev[2];
  if (ev[0] < min_ev) {
    ev[0] = min_ev;
    if (ev[1] < min_ev) ev[1] = min_ev;
    double Einv[9], ED[9];
    inverse3_cofactor(E, Einv);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) ED[r * 3 + c] = E[r * 3 + c] * ev[c];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        cov[r * 3 + c] = (ED[r * 3 + 0] * Einv[0 * 3 + c] + ED[r * 3 + 1] * Einv[1 * 3 + c]) + ED[r * 3 + 2] * Einv[2 * 3 + c];
  }
  inverse3_cofactor(cov, out->icov);
  double mx = -std::numeric_limits<double>::infinity(), mn = std::numeric_limits<double>::infinity();
  for (int k = 0; k < 9; ++k) { mx = std::max(mx, out->icov[k]); mn = std::min(mn, out->icov[k]); }
  if (mx == std::numeric_limits<double>::infinity() || mn == -std::numeric_limits<double>::infinity())
    out->nr_points = -1;
}

struct NdtAngular {
  float j[8][3];    // rows of j_ang (4th column is zero)
  float h[15][3];   // rows of h_ang (a2,a3,b2,b3,c2,c3,d1,d2,d3,e1,e2,e3,f1,f2,f3)
};

// computeAngleDerivatives (ndt_omp_impl.hpp:288-393): doubles, stored as floats.
inline void ndt_angle_derivatives(const double* p, NdtAngular* a) {
  double cx, cy, cz, sx, sy, sz;
  if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
  if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
  if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
  const double J[8][3] = {
      {(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
      {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
      {(-sy * cz), sy * sz, cy},
      {sx * cy * cz, (-sx * cy * sz), sx * sy},
      {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
      {(-cy * sz), (-cy * cz), 0},
      {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
      {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
  const double H[15][3] = {
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},
      {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},
      {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},
      {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},
      {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},
      {(-cy * cz), (cy * sz), (sy)},
      {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},
      {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},
      {(sy * sz), (sy * cz), 0},
      {(-sx * cy * sz), (-sx * cy * cz), 0},
      {(cx * cy * sz), (cx * cy * cz), 0},
      {(-cy * cz), (cy * sz), 0},
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};
  for (int r = 0; r < 8; ++r) for (int c = 0; c < 3; ++c) a->j[r][c] = (float)J[r][c];
  for (int r = 0; r < 15; ++r) for (int c = 0; c < 3; ++c) a->h[r][c] = (float)H[r][c];
}

// One (point, neighbour voxel) term: computePointDerivatives (float overload) +
// updateDerivatives.  x_orig / x_trans_f: the original and the transformed source point
// (floats).  Adds into grad_pt[6] / hess_pt[36] (row-major) and returns score_inc.
inline double ndt_update_derivatives(const NdtAngular* ang, const NdtGauss* g, const float* x_orig,
                                     const float* x_trans_f, const double* mean, const double* icov,
                                     double* grad_pt, double* hess_pt) {
  // Eigen::Vector3d x(x_pt) then Vector4f x4(x[0], x[1], x[2], 0): float -> double -> float
  const float x4[3] = {x_orig[0], x_orig[1], x_orig[2]};
  float xj[8], xh[15];
  for (int r = 0; r < 8; ++r) xj[r] = (ang->j[r][0] * x4[0] + ang->j[r][1] * x4[1]) + ang->j[r][2] * x4[2];
  for (int r = 0; r < 15; ++r) xh[r] = (ang->h[r][0] * x4[0] + ang->h[r][1] * x4[1]) + ang->h[r][2] * x4[2];
  // point_gradient4: 4x6, rows 0..2 (row 3 is zero)
  float pg[3][6] = {{1, 0, 0, 0, xj[2], xj[5]}, {0, 1, 0, xj[0], xj[3], xj[6]}, {0, 0, 1, xj[1], xj[4], xj[7]}};
  // point_hessian blocks (i,j) for i,j in 3..5 : 3-vectors a..f
  const float va[3] = {0, xh[0], xh[1]}, vb[3] = {0, xh[2], xh[3]}, vc[3] = {0, xh[4], xh[5]};
  const float vd[3] = {xh[6], xh[7], xh[8]}, ve[3] = {xh[9], xh[10], xh[11]}, vf[3] = {xh[12], xh[13], xh[14]};
  const float* ph[3][3] = {{va, vb, vc}, {vb, vd, ve}, {vc, ve, vf}};
  // x_trans -= mean (double), then to float
  float xt[3];
  for (int d = 0; d < 3; ++d) xt[d] = (float)((double)x_trans_f[d] - mean[d]);
  float ci[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) ci[r][c] = (float)icov[r * 3 + c];
  float xc[3];   // x_trans4 * c_inv4
  for (int c = 0; c < 3; ++c) xc[c] = (xt[0] * ci[0][c] + xt[1] * ci[1][c]) + xt[2] * ci[2][c];
  const float q = (xt[0] * xc[0] + xt[1] * xc[1]) + xt[2] * xc[2];
  const float d2f = (float)g->d2;
  float e = std::exp(((-d2f) * q) * 0.5f);
  const float score_inc = (float)(-g->d1 * (double)e);
  e = d2f * e;
  if (e > 1 || e < 0 || e != e) return 0.0;
  e = (float)((double)e * g->d1);
  float cg[3][6];   // c_inv4 * point_gradient4
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 6; ++c) cg[r][c] = (ci[r][0] * pg[0][c] + ci[r][1] * pg[1][c]) + ci[r][2] * pg[2][c];
  float xg[6];      // x_trans4 * (c_inv4 * point_gradient4)
  for (int c = 0; c < 6; ++c) xg[c] = (xt[0] * cg[0][c] + xt[1] * cg[1][c]) + xt[2] * cg[2][c];
  for (int c = 0; c < 6; ++c) grad_pt[c] += (double)(e * xg[c]);
  float gg[6][6];   // point_gradient4^T * (c_inv4 * point_gradient4)
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) gg[a][b] = (pg[0][a] * cg[0][b] + pg[1][a] * cg[1][b]) + pg[2][a] * cg[2][b];
  for (int i = 0; i < 6; ++i) {
    float hrow[6] = {0, 0, 0, 0, 0, 0};   // (x_trans4 * c_inv4) * point_hessian.block<4,6>(4i, 0)
    if (i >= 3)
      for (int j = 3; j < 6; ++j) {
        const float* v = ph[i - 3][j - 3];
        hrow[j] = (xc[0] * v[0] + xc[1] * v[1]) + xc[2] * v[2];
      }
    for (int j = 0; j < 6; ++j)
      hess_pt[i * 6 + j] += (double)(e * ((((-d2f) * xg[i]) * xg[j] + hrow[j]) + gg[j][i]));
  }
  return (double)score_inc;
}

// ---- stock pcl::NormalDistributionsTransform (PCL 1.8 ndt.hpp): the same derivatives with
// double Eigen matrices (the in-tree double overloads ndt_omp_impl.hpp:441-481 and
// updateHessian :596-629 are that code).  Used by NdtWithGicp (ndt_gicp.h:62-68).
struct NdtAngularD {
  double j[8][3];
  double h[15][3];
};
inline void ndt_angle_derivatives_f64(const double* p, NdtAngularD* a) {
  NdtAngular tmp;   // same formulas; recompute in double instead of copying the float casts
  (void)tmp;
  double cx, cy, cz, sx, sy, sz;
  if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
  if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
  if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
  const double J[8][3] = {
      {(-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy)},
      {(cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy)},
      {(-sy * cz), sy * sz, cy},
      {sx * cy * cz, (-sx * cy * sz), sx * sy},
      {(-cx * cy * cz), cx * cy * sz, (-cx * sy)},
      {(-cy * sz), (-cy * cz), 0},
      {(cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0},
      {(sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0}};
  const double H[15][3] = {
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy},
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy)},
      {(cx * cy * cz), (-cx * cy * sz), (cx * sy)},
      {(sx * cy * cz), (-sx * cy * sz), (sx * sy)},
      {(-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0},
      {(cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0},
      {(-cy * cz), (cy * sz), (sy)},
      {(-sx * sy * cz), (sx * sy * sz), (sx * cy)},
      {(cx * sy * cz), (-cx * sy * sz), (-cx * cy)},
      {(sy * sz), (sy * cz), 0},
      {(-sx * cy * sz), (-sx * cy * cz), 0},
      {(cx * cy * sz), (cx * cy * cz), 0},
      {(-cy * cz), (cy * sz), 0},
      {(-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0},
      {(-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0}};
  for (int r = 0; r < 8; ++r) for (int c = 0; c < 3; ++c) a->j[r][c] = J[r][c];
  for (int r = 0; r < 15; ++r) for (int c = 0; c < 3; ++c) a->h[r][c] = H[r][c];
}

inline double ndt_update_derivatives_f64(const NdtAngularD* ang, const NdtGauss* g, const float* x_orig,
                                         const float* x_trans_f, const double* mean, const double* icov,
                                         double* grad_pt, double* hess_pt) {
  const double x[3] = {(double)x_orig[0], (double)x_orig[1], (double)x_orig[2]};
  double xj[8], xh[15];
  for (int r = 0; r < 8; ++r) xj[r] = (x[0] * ang->j[r][0] + x[1] * ang->j[r][1]) + x[2] * ang->j[r][2];
  for (int r = 0; r < 15; ++r) xh[r] = (x[0] * ang->h[r][0] + x[1] * ang->h[r][1]) + x[2] * ang->h[r][2];
  const double pg[3][6] = {{1, 0, 0, 0, xj[2], xj[5]}, {0, 1, 0, xj[0], xj[3], xj[6]}, {0, 0, 1, xj[1], xj[4], xj[7]}};
  const double va[3] = {0, xh[0], xh[1]}, vb[3] = {0, xh[2], xh[3]}, vc[3] = {0, xh[4], xh[5]};
  const double vd[3] = {xh[6], xh[7], xh[8]}, ve[3] = {xh[9], xh[10], xh[11]}, vf[3] = {xh[12], xh[13], xh[14]};
  const double* ph[3][3] = {{va, vb, vc}, {vb, vd, ve}, {vc, ve, vf}};
  double xt[3];
  for (int d = 0; d < 3; ++d) xt[d] = (double)x_trans_f[d] - mean[d];
  double cx[3];   // c_inv * x_trans
  for (int r = 0; r < 3; ++r) cx[r] = (icov[r * 3] * xt[0] + icov[r * 3 + 1] * xt[1]) + icov[r * 3 + 2] * xt[2];
  double e = std::exp(-g->d2 * ((xt[0] * cx[0] + xt[1] * cx[1]) + xt[2] * cx[2]) / 2);
  const double score_inc = -g->d1 * e;
  e = g->d2 * e;
  if (e > 1 || e < 0 || e != e) return 0.0;
  e *= g->d1;
  double cdp[6][3], xdot[6];   // c_inv * point_gradient.col(i), x_trans . that
  for (int i = 0; i < 6; ++i) {
    for (int r = 0; r < 3; ++r) cdp[i][r] = (icov[r * 3] * pg[0][i] + icov[r * 3 + 1] * pg[1][i]) + icov[r * 3 + 2] * pg[2][i];
    xdot[i] = (xt[0] * cdp[i][0] + xt[1] * cdp[i][1]) + xt[2] * cdp[i][2];
  }
  for (int i = 0; i < 6; ++i) {
    grad_pt[i] += xdot[i] * e;
    for (int j = 0; j < 6; ++j) {
      double hterm = 0.0;   // x_trans . (c_inv * point_hessian.block<3,1>(3i, j))
      if (i >= 3 && j >= 3) {
        const double* v = ph[i - 3][j - 3];
        double cv[3];
        for (int r = 0; r < 3; ++r) cv[r] = (icov[r * 3] * v[0] + icov[r * 3 + 1] * v[1]) + icov[r * 3 + 2] * v[2];
        hterm = (xt[0] * cv[0] + xt[1] * cv[1]) + xt[2] * cv[2];
      }
      const double gdot = (pg[0][j] * cdp[i][0] + pg[1][j] * cdp[i][1]) + pg[2][j] * cdp[i][2];
      hess_pt[i * 6 + j] += e * ((-g->d2 * xdot[i] * xdot[j] + hterm) + gdot);
    }
  }
  return score_inc;
}

// Translation(p0..2) * AngleAxis(p3, X) * AngleAxis(p4, Y) * AngleAxis(p5, Z), single
// precision (ndt_omp_impl.hpp:146-149, 809-812).  T: 4x4 column-major floats.
inline void axis_rotation_f(float angle, int axis, float* R) {   // AngleAxis::toRotationMatrix
  const float s = std::sin(angle), c = std::cos(angle);
  float ax[3] = {0, 0, 0};
  ax[axis] = 1.0f;
  const float sa[3] = {s * ax[0], s * ax[1], s * ax[2]};
  const float ca[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float t;
  t = ca[0] * ax[1]; R[1] = t - sa[2]; R[3] = t + sa[2];
  t = ca[0] * ax[2]; R[2] = t + sa[1]; R[6] = t - sa[1];
  t = ca[1] * ax[2]; R[5] = t - sa[0]; R[7] = t + sa[0];
  R[0] = ca[0] * ax[0] + c; R[4] = ca[1] * ax[1] + c; R[8] = ca[2] * ax[2] + c;
}
inline void mul3f(const float* A, const float* B, float* C) {   // row-major 3x3
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = (A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c]) + A[r * 3 + 2] * B[6 + c];
}
inline void ndt_transform_from_p(const double* p, float* T) {
  float Rx[9], Ry[9], Rz[9], Rxy[9], R[9];
  axis_rotation_f((float)p[3], 0, Rx);
  axis_rotation_f((float)p[4], 1, Ry);
  axis_rotation_f((float)p[5], 2, Rz);
  mul3f(Rx, Ry, Rxy);
  mul3f(Rxy, Rz, R);
  for (int i = 0; i < 16; ++i) T[i] = 0.0f;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[r + 4 * c] = R[r * 3 + c];
    T[12 + r] = (float)p[r];
  }
  T[15] = 1.0f;
}

// translation + rotation().eulerAngles(0, 1, 2) in single precision (ndt_omp_impl.hpp:103-111;
// Eigen 3.3 EulerAngles.h)
inline void ndt_p_from_transform(const float* T, double* p) {
  const float m00 = T[0], m01 = T[4], m02 = T[8], m10 = T[1], m11 = T[5], m12 = T[9], m20 = T[2], m21 = T[6], m22 = T[10];
  (void)m10; (void)m20;
  const float pi = 3.14159265358979323846f;
  float r0 = std::atan2(m12, m22);
  const float c2 = std::sqrt(m00 * m00 + m01 * m01);
  float r1;
  if (r0 > 0.0f) {   // even permutation (0,1,2): odd = 0
    r0 -= pi;
    r1 = std::atan2(-m02, -c2);
  } else {
    r1 = std::atan2(-m02, c2);
  }
  const float s1 = std::sin(r0), c1 = std::cos(r0);
  const float r2 = std::atan2(s1 * T[2 + 4 * 0] - c1 * T[1 + 4 * 0], c1 * m11 - s1 * m21);
  p[0] = T[12]; p[1] = T[13]; p[2] = T[14];
  p[3] = -r0; p[4] = -r1; p[5] = -r2;
}

// pcl::transformPointCloud for one point, single precision, left to right
inline void ndt_transform_point(const float* T, const float* in, float* out) {
  for (int r = 0; r < 3; ++r) out[r] = ((T[r] * in[0] + T[r + 4] * in[1]) + T[r + 8] * in[2]) + T[r + 12];
}

}  // namespace sm_oracle

#endif  // ORACLE_NDT_MATH_H_
