// ORACLE — TEST INFRASTRUCTURE ONLY (see sm_oracle.h).  PARITY UNPINNED.
//
// CPU restatement of registrator::Ndt (registrators/ndt.cc:28-64) = the vendored pclomp NDT
// (registrators/pclomp/ndt_omp_impl.hpp, voxel_grid_covariance_omp_impl.hpp) as configured by
// the reference: resolution 1.0, KDTREE (centroid radius) neighbour search, step 0.1,
// outlier ratio 0.55, transformation epsilon 0.1, 35 iterations.  The pieces that live in
// stock PCL (pcl::Registration::align / getFitnessScore, transformPointCloud, FLANN searches)
// are restated from their published behaviour; the reference's call sites are cited.
//
// Facts of the reference code that this restatement keeps on purpose:
//  * Leaf::cov_ starts as IDENTITY and x*x^T is added to it (voxel_grid_covariance_omp.h:96-106,
//    _impl.hpp:233), so every covariance carries an extra I/n.
//  * eigenvalue(1) is only inflated inside the branch that inflates eigenvalue(0) (_impl.hpp:345-353).
//  * leaves with >= 6 points whose eigenvalues fail the test keep their centroid in the search
//    cloud with icov = 0 (_impl.hpp:297-341): they add -d1 to the score and nothing else.
//  * computeStepLengthMT starts with interval_converged = (step_max - step_min) > 0 == true
//    (ndt_omp_impl.hpp:802), so the More-Thuente loop never runs: each outer iteration is ONE
//    derivative evaluation at a step clamped to [0.05, 0.1], and computeHessian is never called.
//  * the per-point derivative math is single precision (Matrix<float,4,6>), sums are double.
#include "sm_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "linalg.h"
#include "ndt_math.h"

namespace sm_oracle {

int NdtAlignCore(const float* source, int64_t ns, const float* target, int64_t nt, const double* guess,
                 const sm_oracle_ndt_options* opt, bool f64_math, double* result, double* fitness,
                 int32_t* iterations, double* trans_probability, double* mean_neighbors);
// implemented in sm_oracle.cc
int ExactNn1Float(const float* target_xyz, int64_t nt, const float* query_xyz, int64_t nq,
                  int32_t* ids_out);

namespace {

struct Leaf {
  int nr_points = 0;
  double sum[3] = {0, 0, 0};
  double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // identity on purpose (see header)
  float centroid[3] = {0, 0, 0};
  double mean[3] = {0, 0, 0};
  double icov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool searchable = false;
};

struct VoxelGrid {
  float inv_leaf = 1.0f;
  int min_b[3], div_b[3], mul[3];
  std::map<int, Leaf> leaves;
  std::vector<int> centroid_leaf;     // searchable leaves in ascending idx order

  void Build(const float* pts, int64_t n, float resolution) {
    inv_leaf = 1.0f / resolution;
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (int64_t i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], pts[3 * i + d]); mx[d] = std::max(mx[d], pts[3 * i + d]); }
    int max_b[3];
    for (int d = 0; d < 3; ++d) {   // _impl.hpp:88-97
      min_b[d] = (int)std::floor(mn[d] * inv_leaf);
      max_b[d] = (int)std::floor(mx[d] * inv_leaf);
      div_b[d] = max_b[d] - min_b[d] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    leaves.clear();
    for (int64_t i = 0; i < n; ++i) {   // _impl.hpp:209-263
      const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
      const int idx = ndt_voxel_index(x, y, z, inv_leaf, min_b, mul);
      Leaf& leaf = leaves[idx];
      const double p[3] = {(double)x, (double)y, (double)z};
      for (int d = 0; d < 3; ++d) leaf.sum[d] += p[d];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) leaf.cov[r * 3 + c] += p[r] * p[c];
      leaf.centroid[0] += x; leaf.centroid[1] += y; leaf.centroid[2] += z;
      ++leaf.nr_points;
    }
    centroid_leaf.clear();
    for (auto& kv : leaves) {   // _impl.hpp:283-366
      Leaf& leaf = kv.second;
      NdtLeafOut out;
      ndt_finalize_leaf(leaf.nr_points, leaf.sum, leaf.cov, leaf.centroid, 6, 0.01, &out);
      for (int d = 0; d < 3; ++d) { leaf.mean[d] = out.mean[d]; leaf.centroid[d] = out.centroid[d]; }
      for (int k = 0; k < 9; ++k) leaf.icov[k] = out.icov[k];
      leaf.searchable = out.searchable != 0;
      leaf.nr_points = out.nr_points;
      if (leaf.searchable) centroid_leaf.push_back(kv.first);
    }
  }

  // VoxelGridCovariance::radiusSearch (voxel_grid_covariance_omp.h:471-499): centroids with
  // squared float distance < radius^2, sorted by distance (FLANN sorted = true).  All such
  // centroids lie in the 3x3x3 voxel block around the query's voxel when radius <= leaf size.
  int Neighbors(float qx, float qy, float qz, float radius, const Leaf** out, int cap) const {
    struct Cand { float d; int idx; const Leaf* leaf; };
    Cand cand[27];
    int nc = 0;
    const int i0 = (int)(std::floor(qx * inv_leaf) - (float)min_b[0]);
    const int i1 = (int)(std::floor(qy * inv_leaf) - (float)min_b[1]);
    const int i2 = (int)(std::floor(qz * inv_leaf) - (float)min_b[2]);
    const float r2 = radius * radius;
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int a = i0 + dx, b = i1 + dy, c = i2 + dz;
          if (a < 0 || b < 0 || c < 0 || a >= div_b[0] || b >= div_b[1] || c >= div_b[2]) continue;
          const int idx = a * mul[0] + b * mul[1] + c * mul[2];
          auto it = leaves.find(idx);
          if (it == leaves.end() || !it->second.searchable) continue;
          const float d = ndt_dist2f(qx, qy, qz, it->second.centroid);
          if (d < r2) cand[nc++] = Cand{d, idx, &it->second};
        }
    std::sort(cand, cand + nc, [](const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.idx < b.idx); });
    const int k = std::min(nc, cap);
    for (int i = 0; i < k; ++i) out[i] = cand[i].leaf;
    return k;
  }
};

// computeDerivatives (ndt_omp_impl.hpp:180-284), serial, points in index order.
void ComputeDerivatives(const VoxelGrid& grid, const float* src, const float* trans, int64_t n,
                        const double p[6], const NdtGauss& g, float resolution, double* score_out,
                        double grad[6], double hess[36], double* mean_neighbors, bool f64_math = false) {
  NdtAngular ang;
  NdtAngularD angd;
  ndt_angle_derivatives(p, &ang);
  ndt_angle_derivatives_f64(p, &angd);
  double score = 0.0;
  for (int i = 0; i < 6; ++i) grad[i] = 0.0;
  for (int i = 0; i < 36; ++i) hess[i] = 0.0;
  int64_t total_nb = 0;
  for (int64_t i = 0; i < n; ++i) {
    const Leaf* nb[27];
    const int k = grid.Neighbors(trans[3 * i], trans[3 * i + 1], trans[3 * i + 2], resolution, nb, 27);
    total_nb += k;
    double score_pt = 0.0, grad_pt[6] = {0, 0, 0, 0, 0, 0}, hess_pt[36];
    for (int q = 0; q < 36; ++q) hess_pt[q] = 0.0;
    for (int j = 0; j < k; ++j)
      score_pt += f64_math
                      ? ndt_update_derivatives_f64(&angd, &g, src + 3 * i, trans + 3 * i, nb[j]->mean, nb[j]->icov,
                                                   grad_pt, hess_pt)
                      : ndt_update_derivatives(&ang, &g, src + 3 * i, trans + 3 * i, nb[j]->mean, nb[j]->icov,
                                               grad_pt, hess_pt);
    score += score_pt;
    for (int q = 0; q < 6; ++q) grad[q] += grad_pt[q];
    for (int q = 0; q < 36; ++q) hess[q] += hess_pt[q];
  }
  *score_out = score;
  if (mean_neighbors) *mean_neighbors = n > 0 ? (double)total_nb / (double)n : 0.0;
}

// One-sided Jacobi SVD of a 6x6 (row-major) matrix; solves A x = b like
// Eigen::JacobiSVD(ComputeFullU|ComputeFullV).solve (ndt_omp_impl.hpp:127-129).
void SvdSolve6(const double* A, const double* b, double* x) {
  double U[36], V[36];
  for (int i = 0; i < 36; ++i) { U[i] = A[i]; V[i] = 0.0; }
  for (int i = 0; i < 6; ++i) V[i * 6 + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 6; ++p)
      for (int q = p + 1; q < 6; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 6; ++k) { alpha += U[k * 6 + p] * U[k * 6 + p]; beta += U[k * 6 + q] * U[k * 6 + q]; gamma += U[k * 6 + p] * U[k * 6 + q]; }
        if (gamma == 0.0) continue;
        off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta + std::numeric_limits<double>::min()));
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 6; ++k) {
          const double up = U[k * 6 + p], uq = U[k * 6 + q];
          U[k * 6 + p] = c * up - s * uq; U[k * 6 + q] = s * up + c * uq;
          const double vp = V[k * 6 + p], vq = V[k * 6 + q];
          V[k * 6 + p] = c * vp - s * vq; V[k * 6 + q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  double sv[6], svmax = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += U[k * 6 + j] * U[k * 6 + j];
    sv[j] = std::sqrt(s);
    svmax = std::max(svmax, sv[j]);
  }
  const double thr = std::max(svmax * 6.0 * std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::min());
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    if (!(sv[j] > thr)) continue;
    double dot = 0.0;   // u_j^T b / sigma_j, u_j = U[:,j]/sigma_j
    for (int k = 0; k < 6; ++k) dot += U[k * 6 + j] * b[k];
    dot /= (sv[j] * sv[j]);
    for (int i = 0; i < 6; ++i) x[i] += V[i * 6 + j] * dot;
  }
}

}  // namespace
}  // namespace sm_oracle

using namespace sm_oracle;

extern "C" {

int64_t sm_oracle_ndt_voxels(const float* target, int64_t n, float resolution, int64_t capacity,
                             int32_t* idx_out, int32_t* npts_out, double* mean_out,
                             double* icov_out, float* centroid_out, int32_t* searchable_out) {
  VoxelGrid grid;
  grid.Build(target, n, resolution);
  int64_t k = 0;
  for (const auto& kv : grid.leaves) {
    if (k < capacity) {
      idx_out[k] = kv.first; npts_out[k] = kv.second.nr_points;
      for (int d = 0; d < 3; ++d) { mean_out[3 * k + d] = kv.second.mean[d]; centroid_out[3 * k + d] = kv.second.centroid[d]; }
      for (int q = 0; q < 9; ++q) icov_out[9 * k + q] = kv.second.icov[q];
      searchable_out[k] = kv.second.searchable ? 1 : 0;
    }
    ++k;
  }
  return k;
}

int sm_oracle_ndt_derivatives(const float* source, int64_t ns, const float* target, int64_t nt,
                              const sm_oracle_ndt_options* opt, const double* p, double* score,
                              double* grad6, double* hess36, double* mean_neighbors) {
  VoxelGrid grid;
  grid.Build(target, nt, opt->resolution);
  NdtGauss g;
  ndt_gauss_constants(opt->outlier_ratio, opt->resolution, &g);
  float T[16];
  ndt_transform_from_p(p, T);
  std::vector<float> trans((size_t)(3 * ns));
  for (int64_t i = 0; i < ns; ++i) ndt_transform_point(T, source + 3 * i, &trans[(size_t)(3 * i)]);
  ComputeDerivatives(grid, source, trans.data(), ns, p, g, opt->resolution, score, grad6, hess36, mean_neighbors);
  return 0;
}

// Ndt::Align (ndt.cc:38-64): pcl::Registration::align -> computeTransformation
// (ndt_omp_impl.hpp:81-171) -> getFitnessScore.  guess/result: 4x4 column-major doubles.
int sm_oracle_ndt_align(const float* source, int64_t ns, const float* target, int64_t nt,
                        const double* guess, const sm_oracle_ndt_options* opt, double* result,
                        double* fitness, int32_t* iterations, double* trans_probability,
                        double* mean_neighbors) {
  return sm_oracle::NdtAlignCore(source, ns, target, nt, guess, opt, false, result, fitness, iterations,
                                 trans_probability, mean_neighbors);
}

}  // extern "C"

namespace sm_oracle {
// f64_math = false: the vendored pclomp NDT (single-precision term math);
// f64_math = true : stock pcl::NormalDistributionsTransform (double Eigen matrices), the class
//                   NdtWithGicp uses (registrators/ndt_gicp.h:62-68) — same algorithm otherwise.
int NdtAlignCore(const float* source, int64_t ns, const float* target, int64_t nt, const double* guess,
                 const sm_oracle_ndt_options* opt, bool f64_math, double* result, double* fitness,
                 int32_t* iterations, double* trans_probability, double* mean_neighbors) {
  if (ns <= 0 || nt <= 0) return 0;   // ndt.cc:40-42 returns false when a cloud is missing
  VoxelGrid grid;
  grid.Build(target, nt, opt->resolution);
  NdtGauss g;
  ndt_gauss_constants(opt->outlier_ratio, opt->resolution, &g);
  float final_T[16], guess_f[16];
  bool guess_is_identity = true;
  for (int i = 0; i < 16; ++i) {
    guess_f[i] = (float)guess[i];                       // guess.cast<float>() (ndt.cc:57)
    final_T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    if (guess_f[i] != final_T[i]) guess_is_identity = false;
  }
  std::vector<float> trans(source, source + 3 * ns);     // output = input (PCL align)
  if (!guess_is_identity) {                              // ndt_omp_impl.hpp:95-101
    std::memcpy(final_T, guess_f, sizeof(final_T));
    for (int64_t i = 0; i < ns; ++i) ndt_transform_point(guess_f, source + 3 * i, &trans[(size_t)(3 * i)]);
  }
  double p[6];
  ndt_p_from_transform(final_T, p);                      // :103-111
  double score = 0.0, grad[6], hess[36], nbm = 0.0, nb_sum = 0.0;
  int evals = 0;
  ComputeDerivatives(grid, source, trans.data(), ns, p, g, opt->resolution, &score, grad, hess, &nbm, f64_math);
  nb_sum += nbm; ++evals;
  int nr_iterations = 0;
  bool converged = false;
  while (!converged) {
    double neg_grad[6], delta[6];
    for (int i = 0; i < 6; ++i) neg_grad[i] = -grad[i];
    SvdSolve6(hess, neg_grad, delta);                    // :127-129
    double norm = 0.0;
    for (int i = 0; i < 6; ++i) norm += delta[i] * delta[i];
    norm = std::sqrt(norm);
    if (norm == 0.0 || norm != norm) break;              // :134-139
    for (int i = 0; i < 6; ++i) delta[i] /= norm;
    // computeStepLengthMT with the loop that never runs (:757-916)
    double d_phi_0 = 0.0;
    for (int i = 0; i < 6; ++i) d_phi_0 += grad[i] * delta[i];
    d_phi_0 = -d_phi_0;
    double a_t = 0.0;
    if (d_phi_0 >= 0.0) {
      if (d_phi_0 == 0.0) { a_t = 0.0; goto step_done; }
      for (int i = 0; i < 6; ++i) delta[i] = -delta[i];  // reverse direction (:777-781)
    }
    a_t = std::max(std::min(norm, opt->step_size), opt->transformation_epsilon / 2.0);
    {
      double x_t[6];
      for (int i = 0; i < 6; ++i) x_t[i] = p[i] + delta[i] * a_t;
      ndt_transform_from_p(x_t, final_T);                // :809-812
      for (int64_t i = 0; i < ns; ++i) ndt_transform_point(final_T, source + 3 * i, &trans[(size_t)(3 * i)]);
      ComputeDerivatives(grid, source, trans.data(), ns, x_t, g, opt->resolution, &score, grad, hess, &nbm, f64_math);
      nb_sum += nbm; ++evals;
    }
  step_done:
    for (int i = 0; i < 6; ++i) p[i] += delta[i] * a_t;  // delta_p *= norm; p += delta_p (:143,152)
    if (nr_iterations > opt->max_iterations ||
        (nr_iterations && std::fabs(a_t) < opt->transformation_epsilon))
      converged = true;                                   // :158-162
    ++nr_iterations;
  }
  *trans_probability = score / (double)ns;               // :170
  *iterations = nr_iterations;
  if (mean_neighbors) *mean_neighbors = evals ? nb_sum / evals : 0.0;
  // pcl::Registration::getFitnessScore (ndt.cc:60): mean squared distance (float) to the
  // exact nearest neighbour in the FULL target, over the transformed source.
  std::vector<float> moved((size_t)(3 * ns));
  for (int64_t i = 0; i < ns; ++i) ndt_transform_point(final_T, source + 3 * i, &moved[(size_t)(3 * i)]);
  std::vector<int32_t> ids((size_t)ns);
  ExactNn1Float(target, nt, moved.data(), ns, ids.data());
  double fit = 0.0;
  int64_t nr = 0;
  for (int64_t i = 0; i < ns; ++i) {
    if (ids[(size_t)i] < 0) continue;
    fit += (double)ndt_dist2f(moved[(size_t)(3 * i)], moved[(size_t)(3 * i + 1)], moved[(size_t)(3 * i + 2)],
                              target + 3 * (int64_t)ids[(size_t)i]);
    ++nr;
  }
  *fitness = nr > 0 ? fit / (double)nr : std::numeric_limits<double>::max();
  for (int i = 0; i < 16; ++i) result[i] = (double)final_T[i];   // .cast<double>() (ndt.cc:61)
  return 1;
}

}  // namespace sm_oracle
