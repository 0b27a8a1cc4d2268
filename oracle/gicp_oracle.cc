// ORACLE — TEST INFRASTRUCTURE ONLY (see sm_oracle.h).  PARITY UNPINNED.
//
// CPU restatement of registrator::NdtWithGicp (registrators/ndt_gicp.cc:28-112).  Everything it
// calls lives in stock PCL (distro-pinned 1.8/1.10, not vendored):
//   * pcl::ApproximateVoxelGrid            (filters/impl/approximate_voxel_grid.hpp) — restated
//   * pcl::NormalDistributionsTransform    — the double-precision twin of the vendored pclomp
//                                            code (oracle/ndt_oracle.cc, f64_math = true)
//   * pcl::GeneralizedIterativeClosestPoint — the vendored registrators/pclomp/gicp_omp_impl.hpp
//                                            is the in-tree statement of that math and is followed
//                                            line by line (:59-131 covariances, :255-377 cost and
//                                            gradient, :381-514 outer loop, :516-527 applyState)
//   * pcl/registration/bfgs.h              — a port of GSL's vector_bfgs2 minimiser with
//                                            Fletcher's line search (GSL multimin/linear_minimize.c);
//                                            restated from the published GSL algorithm with the
//                                            parameters gicp sets (gicp_omp_impl.hpp:218-224)
#include "sm_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#include "linalg.h"
#include "ndt_math.h"

namespace sm_oracle {

int NdtAlignCore(const float* source, int64_t ns, const float* target, int64_t nt, const double* guess,
                 const sm_oracle_ndt_options* opt, bool f64_math, double* result, double* fitness,
                 int32_t* iterations, double* trans_probability, double* mean_neighbors);
int ExactNn1Float(const float* target_xyz, int64_t nt, const float* query_xyz, int64_t nq, int32_t* ids_out);
int ExactKnnFloat(const float* target_xyz, int64_t nt, const float* query_xyz, int64_t nq, int k,
                  int32_t* ids_out, double* d2_out);

// ---------------------------------------------------------------------------------------------
// pcl::ApproximateVoxelGrid<PointXYZ>::applyFilter: one pass, a 512-entry hash "history";
// a slot is flushed (centroid emitted) when a point of a different voxel hashes into it, and all
// occupied slots are flushed in slot order at the end.  Output order = flush order.
// ---------------------------------------------------------------------------------------------
void ApproxVoxelGrid(const float* pts, int64_t n, float leaf, std::vector<float>* out) {
  struct He { int ix, iy, iz, count; float c[3]; };
  const int kHist = 512;
  std::vector<He> hist((size_t)kHist, He{0, 0, 0, 0, {0, 0, 0}});
  const float inv = 1.0f / leaf;
  out->clear();
  auto flush = [&](He& h) {
    for (int d = 0; d < 3; ++d) out->push_back(h.c[d] / (float)h.count);
  };
  for (int64_t i = 0; i < n; ++i) {
    const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const int ix = (int)std::floor(x * inv), iy = (int)std::floor(y * inv), iz = (int)std::floor(z * inv);
    const unsigned hash = (unsigned)((ix * 7171 + iy * 3079 + iz * 4231) & (kHist - 1));
    He& h = hist[hash];
    if (h.count && (ix != h.ix || iy != h.iy || iz != h.iz)) {
      flush(h);
      h.count = 0; h.c[0] = h.c[1] = h.c[2] = 0.0f;
    }
    h.ix = ix; h.iy = iy; h.iz = iz;
    h.count++;
    h.c[0] += x; h.c[1] += y; h.c[2] += z;
  }
  for (int s = 0; s < kHist; ++s)
    if (hist[(size_t)s].count) flush(hist[(size_t)s]);
}

// ---------------------------------------------------------------------------------------------
// BFGS (GSL vector_bfgs2 + linear_minimize.c, as ported in pcl/registration/bfgs.h)
// ---------------------------------------------------------------------------------------------
namespace bfgs {

enum Status { kSuccess = 0, kRunning = 1, kNoProgress = 2, kNegativeGradientEpsilon = -3 };

struct Params { double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5, step_size = 0.01; int order = 3; };

typedef std::function<void(const double* x, double* f, double* g)> Fdf;   // f and/or g may be null

inline int solve_quadratic(double a, double b, double c, double* x0, double* x1) {
  const double disc = b * b - 4 * a * c;
  if (a == 0) { if (b == 0) return 0; *x0 = -c / b; return 1; }
  if (disc > 0) {
    if (b == 0) { const double r = std::fabs(0.5 * std::sqrt(disc) / a); *x0 = -r; *x1 = r; }
    else {
      const double sgnb = (b > 0 ? 1 : -1);
      const double temp = -0.5 * (b + sgnb * std::sqrt(disc));
      const double r1 = temp / a, r2 = c / temp;
      if (r1 < r2) { *x0 = r1; *x1 = r2; } else { *x0 = r2; *x1 = r1; }
    }
    return 2;
  }
  if (disc == 0) { *x0 = -0.5 * b / a; *x1 = -0.5 * b / a; return 2; }
  return 0;
}
inline double cubic(double c0, double c1, double c2, double c3, double z) { return c0 + z * (c1 + z * (c2 + z * c3)); }
inline void check_extremum(double c0, double c1, double c2, double c3, double z, double* zmin, double* fmin) {
  const double y = cubic(c0, c1, c2, c3, z);
  if (y < *fmin) { *zmin = z; *fmin = y; }
}
inline double interp_quad(double f0, double fp0, double f1, double zl, double zh) {
  const double fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0));
  const double fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0));
  const double c = 2 * (f1 - f0 - fp0);
  double zmin = zl, fmin = fl;
  if (fh < fmin) { zmin = zh; fmin = fh; }
  if (c > 0) {
    const double z = -fp0 / c;
    if (z > zl && z < zh) {
      const double f = f0 + z * (fp0 + z * (f1 - f0 - fp0));
      if (f < fmin) { zmin = z; fmin = f; }
    }
  }
  return zmin;
}
inline double interp_cubic(double f0, double fp0, double f1, double fp1, double zl, double zh) {
  const double eta = 3 * (f1 - f0) - 2 * fp0 - fp1;
  const double xi = fp0 + fp1 - 2 * (f1 - f0);
  const double c0 = f0, c1 = fp0, c2 = eta, c3 = xi;
  double zmin = zl, fmin = cubic(c0, c1, c2, c3, zl), z0 = 0, z1 = 0;
  check_extremum(c0, c1, c2, c3, zh, &zmin, &fmin);
  const int n = solve_quadratic(3 * c3, 2 * c2, c1, &z0, &z1);
  if (n == 2) {
    if (z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
    if (z1 > zl && z1 < zh) check_extremum(c0, c1, c2, c3, z1, &zmin, &fmin);
  } else if (n == 1) {
    if (z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
  }
  return zmin;
}
inline double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin,
                          double xmax, int order) {
  double ymin = (xmin - a) / (b - a), ymax = (xmax - a) / (b - a);
  if (ymin > ymax) std::swap(ymin, ymax);
  double z;
  if (order > 2 && !(fpb != fpb)) z = interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), ymin, ymax);
  else z = interp_quad(fa, fpa * (b - a), fb, ymin, ymax);
  return a + z * (b - a);
}

struct Minimizer {
  static const int N = 6;
  Fdf fdf;
  Params par;
  double x0[N], g0[N], p[N], gradient[N], x_alpha[N], g_alpha[N];
  double f = 0, g0norm = 0, pnorm = 0, fp0 = 0, delta_f = 0;
  // 1-D view along p with a value cache (function_fdf wrapper of GSL / PCL)
  double f_alpha = 0, df_alpha = 0, f_key = 0, df_key = 0, x_key = 0, g_key = 0;

  static double dot(const double* a, const double* b) { double s = 0; for (int i = 0; i < N; ++i) s += a[i] * b[i]; return s; }
  static double norm(const double* a) { return std::sqrt(dot(a, a)); }
  void move_to(double alpha) {
    if (alpha == x_key) return;
    for (int i = 0; i < N; ++i) x_alpha[i] = x0[i] + alpha * p[i];
    x_key = alpha;
  }
  double slope() const { return dot(g_alpha, p); }
  double eval_f(double alpha) {
    if (alpha == f_key) return f_alpha;
    move_to(alpha);
    fdf(x_alpha, &f_alpha, nullptr);
    f_key = alpha;
    return f_alpha;
  }
  double eval_df(double alpha) {
    if (alpha == df_key) return df_alpha;
    move_to(alpha);
    if (alpha != g_key) { fdf(x_alpha, nullptr, g_alpha); g_key = alpha; }
    df_alpha = slope();
    df_key = alpha;
    return df_alpha;
  }
  void eval_fdf(double alpha, double* fo, double* dfo) {
    if (alpha == f_key && alpha == df_key) { *fo = f_alpha; *dfo = df_alpha; return; }
    if (alpha == f_key || alpha == df_key) { *fo = eval_f(alpha); *dfo = eval_df(alpha); return; }
    move_to(alpha);
    fdf(x_alpha, &f_alpha, g_alpha);
    f_key = alpha; g_key = alpha;
    df_alpha = slope(); df_key = alpha;
    *fo = f_alpha; *dfo = df_alpha;
  }
  void update_position(double alpha, double* x, double* fo, double* g) {
    double f_a, df_a;
    eval_fdf(alpha, &f_a, &df_a);
    *fo = f_a;
    for (int i = 0; i < N; ++i) { x[i] = x_alpha[i]; g[i] = g_alpha[i]; }
  }
  void change_direction() {
    for (int i = 0; i < N; ++i) { x_alpha[i] = x0[i]; g_alpha[i] = g0[i]; }
    x_key = 0; f_key = 0; g_key = 0;
    f_alpha = f; df_alpha = slope(); df_key = 0;
  }
  void init(double* x) {
    delta_f = 0;
    fdf(x, &f, gradient);
    for (int i = 0; i < N; ++i) { x0[i] = x[i]; g0[i] = gradient[i]; }
    g0norm = norm(g0);
    for (int i = 0; i < N; ++i) p[i] = gradient[i] * (-1 / g0norm);
    pnorm = norm(p);
    fp0 = -g0norm;
    change_direction();
  }
  int line_search(double alpha1, double* alpha_new) {
    double f0v, fp0v, falpha, falpha_prev, fpalpha = 0, fpalpha_prev, delta, alpha_next;
    double alpha = alpha1, alpha_prev = 0.0;
    double a = 0.0, b = alpha, fa, fb = 0.0, fpa, fpb = 0.0;
    int i = 0;
    eval_fdf(0.0, &f0v, &fp0v);
    falpha_prev = f0v; fpalpha_prev = fp0v;
    fa = f0v; fpa = fp0v;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    while (i++ < 100) {   // bracketing
      falpha = eval_f(alpha);
      if (falpha > f0v + alpha * par.rho * fp0v || falpha >= falpha_prev) {
        a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
        b = alpha; fb = falpha; fpb = nan;
        break;
      }
      fpalpha = eval_df(alpha);
      if (std::fabs(fpalpha) <= -par.sigma * fp0v) { *alpha_new = alpha; return kSuccess; }
      if (fpalpha >= 0) {
        a = alpha; fa = falpha; fpa = fpalpha;
        b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
        break;
      }
      delta = alpha - alpha_prev;
      alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, alpha + delta,
                               alpha + par.tau1 * delta, par.order);
      alpha_prev = alpha; falpha_prev = falpha; fpalpha_prev = fpalpha;
      alpha = alpha_next;
    }
    while (i++ < 100) {   // sectioning
      delta = b - a;
      alpha = interpolate(a, fa, fpa, b, fb, fpb, a + par.tau2 * delta, b - par.tau3 * delta, par.order);
      falpha = eval_f(alpha);
      if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return kNoProgress;
      if (falpha > f0v + par.rho * alpha * fp0v || falpha >= fa) {
        b = alpha; fb = falpha; fpb = nan;
      } else {
        fpalpha = eval_df(alpha);
        if (std::fabs(fpalpha) <= -par.sigma * fp0v) { *alpha_new = alpha; return kSuccess; }
        if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
          b = a; fb = fa; fpb = fpa;
          a = alpha; fa = falpha; fpa = fpalpha;
        } else {
          a = alpha; fa = falpha; fpa = fpalpha;
        }
      }
    }
    return kSuccess;
  }
  int one_step(double* x) {
    double alpha = 0.0, alpha1;
    const double f0v = f;
    if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) return kNoProgress;
    if (delta_f < 0) {
      const double del = std::max(-delta_f, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0v));
      alpha1 = std::min(1.0, 2.0 * del / (-fp0));
    } else {
      alpha1 = std::fabs(par.step_size);
    }
    const int status = line_search(alpha1, &alpha);
    if (status != kSuccess) return status;
    update_position(alpha, x, &f, gradient);
    delta_f = f - f0v;
    double dx0[N], dg0[N];
    for (int i = 0; i < N; ++i) { dx0[i] = x[i] - x0[i]; dg0[i] = gradient[i] - g0[i]; }
    const double dxg = dot(dx0, gradient), dgg = dot(dg0, gradient), dxdg = dot(dx0, dg0), dgnorm = norm(dg0);
    double A = 0, B = 0;
    if (dxdg != 0) { B = dxg / dxdg; A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg; }
    for (int i = 0; i < N; ++i) p[i] = gradient[i] - A * dx0[i] - B * dg0[i];
    for (int i = 0; i < N; ++i) { g0[i] = gradient[i]; x0[i] = x[i]; }
    g0norm = norm(g0);
    pnorm = norm(p);
    const double dir = (dot(p, gradient) > 0) ? -1.0 : 1.0;
    for (int i = 0; i < N; ++i) p[i] *= dir / pnorm;
    pnorm = norm(p);
    fp0 = dot(p, g0);
    change_direction();
    return kSuccess;
  }
  int test_gradient(double epsabs) const {
    if (epsabs < 0) return kNegativeGradientEpsilon;
    return norm(gradient) < epsabs ? kSuccess : kRunning;
  }
};

}  // namespace bfgs

// ---------------------------------------------------------------------------------------------
// GICP
// ---------------------------------------------------------------------------------------------
struct GicpOptions {
  int k_correspondences = 20;          // gicp_omp.h:109
  double gicp_epsilon = 0.001;         // :110
  double rotation_epsilon = 1e-3;      // ndt_gicp.cc:50 (setRotationEpsilon)
  double transformation_epsilon = 5e-4;// :116
  double corr_dist_threshold = 5.0;    // :117
  int max_iterations = 35;             // ndt_gicp.cc:51
  int max_inner_iterations = 20;       // :112
};

// computeCovariances (gicp_omp_impl.hpp:59-131).  covs: n x 9 row-major.
void GicpCovariances(const float* cloud, int64_t n, int k, double eps, std::vector<double>* covs) {
  covs->assign((size_t)(9 * n), 0.0);
  if (k > n) return;   // PCL_ERROR + return: covariances stay as allocated (zeros here)
  std::vector<int32_t> ids((size_t)(n * k));
  std::vector<double> d2((size_t)(n * k));
  ExactKnnFloat(cloud, n, cloud, n, k, ids.data(), d2.data());
  for (int64_t i = 0; i < n; ++i) {
    double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; ++j) {
      const float* pt = cloud + 3 * (int64_t)ids[(size_t)(i * k + j)];
      mean[0] += pt[0]; mean[1] += pt[1]; mean[2] += pt[2];
      cov[0] += pt[0] * pt[0];                         // float products, double sums (:97-105)
      cov[3] += pt[1] * pt[0]; cov[4] += pt[1] * pt[1];
      cov[6] += pt[2] * pt[0]; cov[7] += pt[2] * pt[1]; cov[8] += pt[2] * pt[2];
    }
    for (int d = 0; d < 3; ++d) mean[d] /= (double)k;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b <= a; ++b) {
        cov[a * 3 + b] /= (double)k;
        cov[a * 3 + b] -= mean[a] * mean[b];
        cov[b * 3 + a] = cov[a * 3 + b];
      }
    // JacobiSVD of a symmetric PSD matrix: U = eigenvectors, singular values descending
    double w[3], V[9];
    JacobiEigenSym(cov, 3, w, V);
    int order[3] = {0, 1, 2};
    for (int a = 0; a < 3; ++a)
      for (int b = a + 1; b < 3; ++b)
        if (std::fabs(w[order[b]]) > std::fabs(w[order[a]])) std::swap(order[a], order[b]);
    double* out = &(*covs)[(size_t)(9 * i)];
    for (int q = 0; q < 9; ++q) out[q] = 0.0;
    for (int kk = 0; kk < 3; ++kk) {
      const double v = (kk == 2) ? eps : 1.0;
      const double col[3] = {V[0 * 3 + order[kk]], V[1 * 3 + order[kk]], V[2 * 3 + order[kk]]};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[r * 3 + c] += v * col[r] * col[c];
    }
  }
}

// applyState (gicp_omp_impl.hpp:516-527): t = [Rz*Ry*Rx * t.R | t.t + x(0..2)], single precision
void GicpApplyState(float* t /*col-major 4x4*/, const double* x) {
  float Rx[9], Ry[9], Rz[9], Rzy[9], R[9], old[9], neu[9];
  axis_rotation_f((float)x[3], 0, Rx);
  axis_rotation_f((float)x[4], 1, Ry);
  axis_rotation_f((float)x[5], 2, Rz);
  mul3f(Rz, Ry, Rzy);
  mul3f(Rzy, Rx, R);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) old[r * 3 + c] = t[r + 4 * c];
  mul3f(R, old, neu);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[r + 4 * c] = neu[r * 3 + c];
  t[12] += (float)x[0]; t[13] += (float)x[1]; t[14] += (float)x[2];
}

// computeRDerivative (gicp_omp_impl.hpp:133-183): g[3..5] = <dR/dangle, R>
void GicpRDerivative(const double* x, const double* R /*row-major 3x3*/, double* g) {
  const double phi = x[3], theta = x[4], psi = x[5];
  const double cphi = std::cos(phi), sphi = std::sin(phi), ctheta = std::cos(theta), stheta = std::sin(theta);
  const double cpsi = std::cos(psi), spsi = std::sin(psi);
  const double dPhi[9] = {0, sphi * spsi + cphi * cpsi * stheta, cphi * spsi - cpsi * sphi * stheta,
                          0, -cpsi * sphi + cphi * spsi * stheta, -cphi * cpsi - sphi * spsi * stheta,
                          0, cphi * ctheta, -ctheta * sphi};
  const double dTheta[9] = {-cpsi * stheta, cpsi * ctheta * sphi, cphi * cpsi * ctheta,
                            -spsi * stheta, ctheta * sphi * spsi, cphi * ctheta * spsi,
                            -ctheta, -sphi * stheta, -cphi * stheta};
  const double dPsi[9] = {-ctheta * spsi, -cphi * cpsi - sphi * spsi * stheta, cpsi * sphi - cphi * spsi * stheta,
                          cpsi * ctheta, -cphi * spsi + cpsi * sphi * stheta, sphi * spsi + cphi * cpsi * stheta,
                          0, 0, 0};
  auto inner = [&](const double* a) {   // matricesInnerProd (gicp_omp.h:318-327): sum_ij a(j,i) * R(i,j)
    double r = 0.0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r += a[j * 3 + i] * R[i * 3 + j];
    return r;
  };
  g[3] = inner(dPhi); g[4] = inner(dTheta); g[5] = inner(dPsi);
}

inline void mulpt_f(const float* T, const float* p, float* out) {   // (T * [p;1]) rows 0..2, float
  for (int r = 0; r < 3; ++r) out[r] = ((T[r] * p[0] + T[r + 4] * p[1]) + T[r + 8] * p[2]) + T[r + 12] * 1.0f;
}

// The per-point part of one outer iteration (gicp_omp_impl.hpp:419-463): query = transformation_ * (guess * p),
// exact 1-NN in the target, gate on the squared float distance, M_i = (R C1 R^T + C2)^-1 with
// R = rot(transformation_ * guess) formed in double.  Correspondences come out in ascending source index.
void GicpCorrespond(const float* src, int64_t ns, const float* tgt, int64_t nt, const float* guess,
                    const float* transformation, const std::vector<double>& cov_s, const std::vector<double>& cov_t,
                    double dist_threshold, std::vector<float>* query_io, std::vector<int32_t>* nn_io,
                    std::vector<double>* maha_io, std::vector<int>* si_out, std::vector<int>* ti_out) {
  std::vector<float>& query = *query_io;
  std::vector<int32_t>& nn = *nn_io;
  std::vector<double>& maha = *maha_io;
  std::vector<int>& si = *si_out;
  std::vector<int>& ti = *ti_out;
  si.clear(); ti.clear();
  double TR[16];   // transform_R = transformation_ * guess in double (:423-427)
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += (double)transformation[i + 4 * k] * (double)guess[k + 4 * j];
      TR[i + 4 * j] = s;
    }
  double R[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = TR[r + 4 * c];
  for (int64_t i = 0; i < ns; ++i) {   // query = transformation_ * (guess * p) (:437-439)
    float q1[3];
    mulpt_f(guess, src + 3 * i, q1);
    mulpt_f(transformation, q1, &query[(size_t)(3 * i)]);
  }
  ExactNn1Float(tgt, nt, query.data(), ns, nn.data());
  for (int64_t i = 0; i < ns; ++i) {
    const int j = nn[(size_t)i];
    if (j < 0) continue;
    const float d = ndt_dist2f(query[(size_t)(3 * i)], query[(size_t)(3 * i + 1)], query[(size_t)(3 * i + 2)], tgt + 3 * (int64_t)j);
    if ((double)d < dist_threshold) {   // :448
      const double* C1 = &cov_s[(size_t)(9 * i)];
      const double* C2 = &cov_t[(size_t)(9 * (int64_t)j)];
      double M[9], tmp[9];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        M[r * 3 + c] = (R[r * 3] * C1[c] + R[r * 3 + 1] * C1[3 + c]) + R[r * 3 + 2] * C1[6 + c];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        tmp[r * 3 + c] = ((M[r * 3] * R[c * 3] + M[r * 3 + 1] * R[c * 3 + 1]) + M[r * 3 + 2] * R[c * 3 + 2]) + C2[r * 3 + c];
      inverse3_cofactor(tmp, &maha[(size_t)(9 * i)]);
      si.push_back((int)i); ti.push_back(j);
    }
  }
}

// OptimizationFunctorWithIndices::fdf / operator() / df (gicp_omp_impl.hpp:255-377) at state x over the
// correspondences (si, ti): f = mean res^T M res, g = [2/m sum M res ; <dR/dangle, 2/m sum (base p) (M res)^T>].
void GicpCost(const float* src, const float* tgt, const float* base, const std::vector<double>& maha,
              const std::vector<int>& si, const std::vector<int>& ti, const double* xx, double* f, double* g) {
  const int m = (int)si.size();
  float T[16];
  for (int i = 0; i < 16; ++i) T[i] = base[i];
  GicpApplyState(T, xx);
  double fs = 0.0, gt[3] = {0, 0, 0}, Rm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < m; ++c) {
    const float* ps = src + 3 * (int64_t)si[(size_t)c];
    const float* pt = tgt + 3 * (int64_t)ti[(size_t)c];
    float pp[3], pb[3];
    mulpt_f(T, ps, pp);
    const double res[3] = {(double)(pp[0] - pt[0]), (double)(pp[1] - pt[1]), (double)(pp[2] - pt[2])};
    const double* M = &maha[(size_t)(9 * (int64_t)si[(size_t)c])];
    double temp[3];
    for (int r = 0; r < 3; ++r) temp[r] = (M[r * 3] * res[0] + M[r * 3 + 1] * res[1]) + M[r * 3 + 2] * res[2];
    fs += (res[0] * temp[0] + res[1] * temp[1]) + res[2] * temp[2];
    for (int r = 0; r < 3; ++r) gt[r] += temp[r];
    mulpt_f(base, ps, pb);
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) Rm[r * 3 + cc] += (double)pb[r] * temp[cc];
  }
  if (f) *f = fs / (double)m;
  if (g) {
    for (int r = 0; r < 3; ++r) g[r] = gt[r] * (2.0 / m);
    for (int q = 0; q < 9; ++q) Rm[q] *= 2.0 / m;
    GicpRDerivative(xx, Rm, g);
  }
}

// GeneralizedIterativeClosestPoint::computeTransformation (gicp_omp_impl.hpp:381-514)
int GicpAlign(const float* src, int64_t ns, const float* tgt, int64_t nt, const float* guess /*col-major*/,
              const GicpOptions& o, float* final_T, int* iterations, int* bfgs_evals) {
  std::vector<double> cov_t, cov_s;
  GicpCovariances(tgt, nt, o.k_correspondences, o.gicp_epsilon, &cov_t);
  GicpCovariances(src, ns, o.k_correspondences, o.gicp_epsilon, &cov_s);
  float transformation[16], previous[16], base[16];
  for (int i = 0; i < 16; ++i) { transformation[i] = previous[i] = (i % 5 == 0) ? 1.0f : 0.0f; base[i] = guess[i]; }
  std::vector<double> maha((size_t)(9 * ns));
  for (int64_t i = 0; i < ns; ++i) for (int q = 0; q < 9; ++q) maha[(size_t)(9 * i + q)] = (q % 4 == 0) ? 1.0 : 0.0;
  const double dist_threshold = o.corr_dist_threshold * o.corr_dist_threshold;
  std::vector<float> query((size_t)(3 * ns));
  std::vector<int32_t> nn((size_t)ns);
  int nr_iterations = 0, evals = 0;
  bool converged = false;
  while (!converged) {
    std::vector<int> si, ti;
    GicpCorrespond(src, ns, tgt, nt, guess, transformation, cov_s, cov_t, dist_threshold, &query, &nn, &maha, &si, &ti);
    for (int i = 0; i < 16; ++i) previous[i] = transformation[i];
    const int m = (int)si.size();
    if (m < 4) break;   // NotEnoughPointsException -> caught, loop ends (:470-492)
    // estimateRigidTransformationBFGS (:187-246)
    double x[6] = {(double)transformation[12], (double)transformation[13], (double)transformation[14],
                   std::atan2((double)transformation[2 + 4 * 1], (double)transformation[2 + 4 * 2]),
                   std::asin(-(double)transformation[2 + 4 * 0]),
                   std::atan2((double)transformation[1 + 4 * 0], (double)transformation[0 + 4 * 0])};
    bfgs::Minimizer mz;
    mz.fdf = [&](const double* xx, double* f, double* g) {   // fdf / operator() / df (:255-377)
      ++evals;
      GicpCost(src, tgt, base, maha, si, ti, xx, f, g);
    };
    mz.par = bfgs::Params();   // sigma 0.01, rho 0.01, tau1 9, tau2 0.05, tau3 0.5, order 3 (:218-224)
    mz.init(x);
    int result = bfgs::kRunning, inner = 0;
    do {
      ++inner;
      result = mz.one_step(x);
      if (result) break;
      result = mz.test_gradient(1e-2);
    } while (result == bfgs::kRunning && inner < o.max_inner_iterations);
    if (result == bfgs::kNoProgress || result == bfgs::kSuccess || inner == o.max_inner_iterations) {
      for (int i = 0; i < 16; ++i) transformation[i] = (i % 5 == 0) ? 1.0f : 0.0f;
      GicpApplyState(transformation, x);
    } else {
      break;   // SolverDidntConvergeException -> caught (:489-492)
    }
    double delta = 0.0;   // :474-486
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) {
        const double ratio = (k < 3 && l < 3) ? 1.0 / o.rotation_epsilon : 1.0 / o.transformation_epsilon;
        const double c_delta = ratio * std::fabs((double)previous[k + 4 * l] - (double)transformation[k + 4 * l]);
        if (c_delta > delta) delta = c_delta;
      }
    ++nr_iterations;
    if (nr_iterations >= o.max_iterations || delta < 1) {
      converged = true;
      for (int i = 0; i < 16; ++i) previous[i] = transformation[i];
    }
  }
  // :505-508
  float Rp[9], Rg[9], Rf[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { Rp[r * 3 + c] = previous[r + 4 * c]; Rg[r * 3 + c] = guess[r + 4 * c]; }
  mul3f(Rp, Rg, Rf);
  for (int i = 0; i < 16; ++i) final_T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) final_T[r + 4 * c] = Rf[r * 3 + c];
    final_T[12 + r] = previous[12 + r] + guess[12 + r];
  }
  *iterations = nr_iterations;
  if (bfgs_evals) *bfgs_evals = evals;
  return 1;
}

double FitnessScore(const float* src, int64_t ns, const float* tgt, int64_t nt, const float* T) {
  std::vector<float> moved((size_t)(3 * ns));
  for (int64_t i = 0; i < ns; ++i) ndt_transform_point(T, src + 3 * i, &moved[(size_t)(3 * i)]);
  std::vector<int32_t> ids((size_t)ns);
  ExactNn1Float(tgt, nt, moved.data(), ns, ids.data());
  double fit = 0.0;
  int64_t nr = 0;
  for (int64_t i = 0; i < ns; ++i) {
    if (ids[(size_t)i] < 0) continue;
    fit += (double)ndt_dist2f(moved[(size_t)(3 * i)], moved[(size_t)(3 * i + 1)], moved[(size_t)(3 * i + 2)],
                              tgt + 3 * (int64_t)ids[(size_t)i]);
    ++nr;
  }
  return nr > 0 ? fit / (double)nr : std::numeric_limits<double>::max();
}

}  // namespace sm_oracle

using namespace sm_oracle;

extern "C" {

int64_t sm_oracle_approx_voxel_grid(const float* pts, int64_t n, float leaf, float* out, int64_t capacity) {
  std::vector<float> o;
  ApproxVoxelGrid(pts, n, leaf, &o);
  const int64_t m = (int64_t)o.size() / 3;
  for (int64_t i = 0; i < std::min(m, capacity) * 3; ++i) out[i] = o[(size_t)i];
  return m;
}

int sm_oracle_gicp_covariances(const float* pts, int64_t n, int k, double eps, double* cov_out) {
  std::vector<double> c;
  GicpCovariances(pts, n, k, eps, &c);
  std::memcpy(cov_out, c.data(), sizeof(double) * (size_t)(9 * n));
  return 0;
}

// NdtWithGicp::Align (ndt_gicp.cc:55-112).  Returns 1 / 0 like the reference's bool.
// Test hook for tests/test_oracle_vs_python_restatement.py: the correspondence step and one cost / gradient
// evaluation of GICP for given transformation_ (col-major float 4x4), guess = base_transformation_ and state x.
// maha_out: ns x 9 (row-major, identity where no correspondence), si / ti: capacity ns; returns m.
int64_t sm_oracle_gicp_cost(const float* src, int64_t ns, const float* tgt, int64_t nt, const float* guess,
                            const float* transformation, const double* x, double* f, double* g6, double* maha_out,
                            int32_t* si_out, int32_t* ti_out) {
  GicpOptions o;
  std::vector<double> cov_t, cov_s;
  GicpCovariances(tgt, nt, o.k_correspondences, o.gicp_epsilon, &cov_t);
  GicpCovariances(src, ns, o.k_correspondences, o.gicp_epsilon, &cov_s);
  std::vector<double> maha((size_t)(9 * ns));
  for (int64_t i = 0; i < ns; ++i) for (int q = 0; q < 9; ++q) maha[(size_t)(9 * i + q)] = (q % 4 == 0) ? 1.0 : 0.0;
  std::vector<float> query((size_t)(3 * ns));
  std::vector<int32_t> nn((size_t)ns);
  std::vector<int> si, ti;
  GicpCorrespond(src, ns, tgt, nt, guess, transformation, cov_s, cov_t, o.corr_dist_threshold * o.corr_dist_threshold,
                 &query, &nn, &maha, &si, &ti);
  if (!si.empty()) GicpCost(src, tgt, guess, maha, si, ti, x, f, g6);
  for (size_t q = 0; q < maha.size(); ++q) maha_out[q] = maha[q];
  for (size_t c = 0; c < si.size(); ++c) { si_out[c] = si[c]; ti_out[c] = ti[c]; }
  return (int64_t)si.size();
}

int sm_oracle_ndt_gicp_align(const float* source, int64_t ns, const float* target, int64_t nt,
                             const double* guess, const sm_oracle_ndt_gicp_options* opt, double* result,
                             double* final_score, sm_oracle_ndt_gicp_info* info) {
  std::vector<float> s, t;
  if (opt->using_voxel_filter) {
    ApproxVoxelGrid(source, ns, opt->voxel_resolution, &s);
    ApproxVoxelGrid(target, nt, opt->voxel_resolution, &t);
  } else {
    s.assign(source, source + 3 * ns);
    t.assign(target, target + 3 * nt);
  }
  const int64_t ms = (int64_t)s.size() / 3, mt = (int64_t)t.size() / 3;
  if (info) { std::memset(info, 0, sizeof(*info)); info->n_source_filtered = ms; info->n_target_filtered = mt; }
  float ndt_guess[16];
  for (int i = 0; i < 16; ++i) ndt_guess[i] = (float)guess[i];
  double ndt_score = 0.9;
  if (opt->use_ndt) {   // :82-88, configured at :44-47
    sm_oracle_ndt_options no{1.0f, 0.1, 0.55, 0.01, 35};
    double res[16], tp = 0, nb = 0;
    int32_t it = 0;
    NdtAlignCore(s.data(), ms, t.data(), mt, guess, &no, true, res, &ndt_score, &it, &tp, &nb);
    for (int i = 0; i < 16; ++i) ndt_guess[i] = (float)res[i];
    if (info) { info->ndt_iterations = it; info->ndt_score = ndt_score; }
  }
  double icp_score = 10.0;
  if (ndt_score <= 1.0) {   // :92-103
    float final_T[16];
    int it = 0, evals = 0;
    GicpAlign(s.data(), ms, t.data(), mt, ndt_guess, GicpOptions(), final_T, &it, &evals);
    icp_score = FitnessScore(s.data(), ms, t.data(), mt, final_T);
    *final_score = std::exp(-icp_score);
    for (int i = 0; i < 16; ++i) result[i] = (double)final_T[i];
    if (info) { info->gicp_iterations = it; info->bfgs_evaluations = evals; info->gicp_fitness = icp_score; }
    return 1;
  }
  for (int i = 0; i < 16; ++i) result[i] = guess[i];   // :104-108
  *final_score = std::exp(-icp_score);
  return 0;
}

}  // extern "C"
