// ORACLE — TEST INFRASTRUCTURE ONLY (see sm_oracle.h).  PARITY UNPINNED.
//
// CPU restatement of the registrators/ ICP hot path of StaticMapping:
//   * EigenPointCloud::CalculateNormals      builder/data/cloud_types.cc:73-144,347-368
//   * EigenPointCloud::ApplyTransform        builder/data/cloud_types.cc:288-302
//   * libnabo 1.0.7 KDTREE_LINEAR_HEAP       external; call sites icp_fast.cc:466-467,177-178
//   * IcpFast::Align and helpers             registrators/icp_fast.cc:65-90,100-166,182-324,
//                                            377-405,455-529
// Every function cites the reference lines it follows.  No reference source is copied:
// the reference is Eigen expression code; this is scalar C++ in the same operation order.
#include "sm_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "linalg.h"

namespace sm_oracle {
namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

// ArgMax of cloud_types.cc:41-56 (and libnabo's argMax): first strictly-greater wins,
// starting from maxVal = 0.
inline int ArgMax3(const double v[3]) {
  double max_val = 0.0;
  int max_idx = 0;
  for (int i = 0; i < 3; ++i)
    if (v[i] > max_val) { max_val = v[i]; max_idx = i; }
  return max_idx;
}

struct CoordLess {
  const double* pts;  // 3xN column-major
  int dim;
  int tie_mode;
  bool operator()(int a, int b) const {
    const double ca = pts[3 * (size_t)a + dim], cb = pts[3 * (size_t)b + dim];
    if (tie_mode == 1) return ca < cb;           // reference comparator (CompareDim)
    return ca < cb || (ca == cb && a < b);       // deterministic total order
  }
};

// ---------------------------------------------------------------------------
// Target preparation: normals + one point per leaf.
// ---------------------------------------------------------------------------
struct NormalsBuilder {
  double* points;    // 3xN, modified in place like the reference
  double* normals;   // 3xN
  std::vector<int> indices;
  std::vector<int> indices_to_keep;
  int tie_mode;

  // cloud_types.cc:73-103
  void Leaf(int first, int last) {
    const int count = last - first;
    if (count <= 0) return;
    // tie_mode 0: canonical member order (ascending original index) so that the sums
    // below have one defined rounding; the reference's order is whatever nth_element left.
    if (tie_mode == 0) std::sort(indices.begin() + first, indices.begin() + last);
    double d[7][3];
    double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < count; ++i) {
      const int idx = indices[first + i];
      for (int r = 0; r < 3; ++r) d[i][r] = points[3 * (size_t)idx + r];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[r * 3 + c] += d[i][r] * d[i][c];
    }
    double b[3] = {0, 0, 0};
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < count; ++i) b[r] += d[i][r];
    double mean[3];
    for (int r = 0; r < 3; ++r) mean[r] = b[r] / count;
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int i = 0; i < count; ++i) s += (d[i][r] - mean[r]) * (d[i][c] - mean[c]);
        C[r * 3 + c] = s;
      }
    FullPivQR<3> qr;
    qr.Compute(C);
    if (qr.Rank() + 1 < 3) return;  // :89-91
    double Minv[9];
    PartialPivLuInverse3(M, Minv);  // :93 (dynamic MatrixXd::inverse)
    double nrm[3];
    for (int r = 0; r < 3; ++r)
      nrm[r] = Minv[r * 3 + 0] * b[0] + Minv[r * 3 + 1] * b[1] + Minv[r * 3 + 2] * b[2];
    int k = indices[first];  // :97
    if (tie_mode == 0)
      for (int i = 1; i < count; ++i) k = std::min(k, indices[first + i]);
    indices_to_keep.push_back(k);
    const double sq = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
    const double inv = (sq > 0.0) ? std::sqrt(sq) : 1.0;  // Eigen normalized()
    for (int r = 0; r < 3; ++r) {
      points[3 * (size_t)k + r] = mean[r];
      normals[3 * (size_t)k + r] = (sq > 0.0) ? nrm[r] / inv : nrm[r];
    }
  }

  // cloud_types.cc:105-144
  void Build(int first, int last, const double minv[3], const double maxv[3]) {
    const int count = last - first;
    if (count <= 7) { Leaf(first, last); return; }
    const double ext[3] = {maxv[0] - minv[0], maxv[1] - minv[1], maxv[2] - minv[2]};
    const int cut_dim = ArgMax3(ext);
    const int right_count = count / 2;
    const int left_count = count - right_count;
    std::nth_element(indices.begin() + first, indices.begin() + first + left_count,
                     indices.begin() + last, CoordLess{points, cut_dim, tie_mode});
    const int cut_index = indices[first + left_count];
    const double cut_val = points[3 * (size_t)cut_index + cut_dim];
    double left_max[3] = {maxv[0], maxv[1], maxv[2]};
    left_max[cut_dim] = cut_val;
    double right_min[3] = {minv[0], minv[1], minv[2]};
    right_min[cut_dim] = cut_val;
    Build(first, first + left_count, minv, left_max);
    Build(first + left_count, last, right_min, maxv);
  }
};

// ---------------------------------------------------------------------------
// libnabo 1.0.7  KDTreeUnbalancedPtInLeavesImplicitBoundsStackOpt restatement.
// ---------------------------------------------------------------------------
struct KdNode {
  int dim;          // 0..2 inner, 3 = leaf
  double cut_val;   // inner
  int right_child;  // inner (left child is pos+1)
  int bucket_first; // leaf
  int bucket_count; // leaf
};

struct KdTree {
  const double* cloud = nullptr;  // 3xN column-major
  int64_t n = 0;
  int bucket_size = 8;
  int tie_mode = 0;
  std::vector<KdNode> nodes;
  std::vector<int> buckets;  // point indices in bucket order

  int BuildNodes(std::vector<int>& idx, int first, int last, const double minv[3],
                 const double maxv[3]) {
    const int count = last - first;
    const int pos = (int)nodes.size();
    if (count <= bucket_size) {
      KdNode leaf{3, 0.0, -1, (int)buckets.size(), count};
      if (tie_mode == 0) std::sort(idx.begin() + first, idx.begin() + last);
      for (int i = 0; i < count; ++i) buckets.push_back(idx[first + i]);
      nodes.push_back(leaf);
      return pos;
    }
    const double ext[3] = {maxv[0] - minv[0], maxv[1] - minv[1], maxv[2] - minv[2]};
    const int cut_dim = ArgMax3(ext);
    const int right_count = count / 2;
    const int left_count = count - right_count;
    std::nth_element(idx.begin() + first, idx.begin() + first + left_count,
                     idx.begin() + last, CoordLess{cloud, cut_dim, tie_mode});
    const double cut_val = cloud[3 * (size_t)idx[first + left_count] + cut_dim];
    nodes.push_back(KdNode{cut_dim, cut_val, -1, -1, 0});
    double left_max[3] = {maxv[0], maxv[1], maxv[2]};
    left_max[cut_dim] = cut_val;
    double right_min[3] = {minv[0], minv[1], minv[2]};
    right_min[cut_dim] = cut_val;
    BuildNodes(idx, first, first + left_count, minv, left_max);
    const int right = BuildNodes(idx, first + left_count, last, right_min, maxv);
    nodes[pos].right_child = right;
    return pos;
  }

  void Build(const double* pts, int64_t count, int bucket, int ties) {
    cloud = pts; n = count; bucket_size = bucket; tie_mode = ties;
    nodes.clear(); buckets.clear();
    if (n <= 0) return;
    double minv[3] = {kInf, kInf, kInf}, maxv[3] = {-kInf, -kInf, -kInf};
    std::vector<int> idx((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      idx[(size_t)i] = (int)i;
      for (int r = 0; r < 3; ++r) {
        minv[r] = std::min(minv[r], pts[3 * i + r]);
        maxv[r] = std::max(maxv[r], pts[3 * i + r]);
      }
    }
    nodes.reserve((size_t)(2 * n / std::max(1, bucket / 2) + 16));
    buckets.reserve((size_t)n);
    BuildNodes(idx, 0, (int)n, minv, maxv);
  }

  // recurseKnn, k = 1, allowSelfMatch, maxRadius2 = inf.
  mutable int* visit_counter = nullptr;  // optional: leaves visited by the current query
  void Recurse(const double* q, int n_idx, double rd, double off[3], double max_error2,
               double& head, int& head_idx, int* visits = nullptr) const {
    const KdNode& node = nodes[(size_t)n_idx];
    if (node.dim == 3) {
      if (visits) ++*visits;
      for (int i = 0; i < node.bucket_count; ++i) {
        const int index = buckets[(size_t)(node.bucket_first + i)];
        const double* p = cloud + 3 * (size_t)index;
        double dist = 0.0;
        for (int r = 0; r < 3; ++r) {
          const double diff = q[r] - p[r];
          dist += diff * diff;
        }
        if (dist < head) { head = dist; head_idx = index; }  // strict: first wins
      }
      return;
    }
    const int cd = node.dim;
    const double old_off = off[cd];
    const double new_off = q[cd] - node.cut_val;
    const int near = (new_off > 0.0) ? node.right_child : n_idx + 1;
    const int far = (new_off > 0.0) ? n_idx + 1 : node.right_child;
    Recurse(q, near, rd, off, max_error2, head, head_idx, visits);
    rd += -old_off * old_off + new_off * new_off;
    if (rd <= kInf && rd * max_error2 < head) {
      off[cd] = new_off;
      Recurse(q, far, rd, off, max_error2, head, head_idx, visits);
      off[cd] = old_off;
    }
  }

  // exact k nearest neighbours, result sorted by (squared distance, index) — the unique
  // k smallest under that total order, so the answer does not depend on the visiting order
  struct Cand { double d; int idx; };
  static bool CandLess(const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.idx < b.idx); }
  void RecurseK(const double* q, int n_idx, double rd, double off[3], std::vector<Cand>& best,
                int k) const {
    const KdNode& node = nodes[(size_t)n_idx];
    if (node.dim == 3) {
      for (int i = 0; i < node.bucket_count; ++i) {
        const int index = buckets[(size_t)(node.bucket_first + i)];
        const double* p = cloud + 3 * (size_t)index;
        double dist = 0.0;
        for (int r = 0; r < 3; ++r) { const double diff = q[r] - p[r]; dist += diff * diff; }
        const Cand c{dist, index};
        if ((int)best.size() < k) { best.push_back(c); std::push_heap(best.begin(), best.end(), CandLess); }
        else if (CandLess(c, best.front())) {
          std::pop_heap(best.begin(), best.end(), CandLess);
          best.back() = c;
          std::push_heap(best.begin(), best.end(), CandLess);
        }
      }
      return;
    }
    const int cd = node.dim;
    const double old_off = off[cd];
    const double new_off = q[cd] - node.cut_val;
    const int near = (new_off > 0.0) ? node.right_child : n_idx + 1;
    const int far = (new_off > 0.0) ? n_idx + 1 : node.right_child;
    RecurseK(q, near, rd, off, best, k);
    rd += -old_off * old_off + new_off * new_off;
    if ((int)best.size() < k || rd <= best.front().d) {
      off[cd] = new_off;
      RecurseK(q, far, rd, off, best, k);
      off[cd] = old_off;
    }
  }
  void KnnK(const double* query, int64_t nq, int k, int32_t* ids, double* d2) const {
#pragma omp parallel for schedule(guided, 32)
    for (int64_t i = 0; i < nq; ++i) {
      double off[3] = {0.0, 0.0, 0.0};
      std::vector<Cand> best;
      best.reserve((size_t)k + 1);
      if (!nodes.empty()) RecurseK(query + 3 * i, 0, 0.0, off, best, k);
      std::sort(best.begin(), best.end(), CandLess);
      for (int j = 0; j < k; ++j) {
        ids[i * k + j] = j < (int)best.size() ? best[(size_t)j].idx : -1;
        d2[i * k + j] = j < (int)best.size() ? best[(size_t)j].d : kInf;
      }
    }
  }

  void Knn1(const double* query, int64_t nq, double epsilon, int32_t* ids,
            double* d2, int32_t* visits = nullptr) const {
    const double max_error2 = (1.0 + epsilon) * (1.0 + epsilon);
#pragma omp parallel for schedule(guided, 32)
    for (int64_t i = 0; i < nq; ++i) {
      double off[3] = {0.0, 0.0, 0.0};
      double head = kInf;
      int head_idx = -1;
      int nv = 0;
      if (!nodes.empty()) Recurse(query + 3 * i, 0, 0.0, off, max_error2, head, head_idx, &nv);
      if (visits) visits[i] = nv;
      ids[i] = head_idx;
      d2[i] = head;
    }
  }
};

// ---------------------------------------------------------------------------
// IcpFast pieces.
// ---------------------------------------------------------------------------
// cloud_types.cc:288-296: 4xN homogeneous product, rows 0..2 kept.
inline void ApplyTransform(const double* T, const double* in, double* out, int64_t n) {
  for (int64_t j = 0; j < n; ++j) {
    const double x = in[3 * j], y = in[3 * j + 1], z = in[3 * j + 2];
    for (int i = 0; i < 3; ++i) {
      double s = T[i + 0] * x;
      s = s + T[i + 4] * y;
      s = s + T[i + 8] * z;
      s = s + T[i + 12] * 1.0;
      out[3 * j + i] = s;
    }
  }
}

// icp_fast.cc:86: `const int quantile_index = values.size() * quantile;` with
// quantile = (double)options_.dist_outlier_ratio (a float, icp_fast.h:59).
inline int QuantileIndex(int64_t n, float ratio) {
  return (int)((double)(size_t)n * (double)ratio);
}

// icp_fast.cc:204-254.  A row-major 6x6 (symmetric so layout is moot), path: 0 LLT,
// 1 rank-reduced min-norm, 2 SVD fallback.
inline void SolvePossiblyUnderdetermined(const double* A, const double* b, double* x,
                                         int* path) {
  FullPivQR<6> qr;
  qr.Compute(A);
  if (qr.IsInvertible()) {
    LltSolve(A, b, x, 6);
    if (path) *path = 0;
    return;
  }
  if (path) *path = 1;
  const int rank = qr.Rank();
  // Q1t = Q^T rows 0..rank-1 ; R1 = (Q1t * A * P) rows 0..rank-1   (:220-221)
  double R1[36] = {0};
  for (int r = 0; r < rank; ++r)
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) {
        double qa = 0.0;  // (Q1t*A)(r,k) recomputed per use keeps the code short
        for (int m = 0; m < 6; ++m) qa += qr.Qt[r * 6 + m] * A[m * 6 + k];
        s += qa * ((qr.perm[c] == k) ? 1.0 : 0.0);
      }
      R1[r * 6 + c] = s;
    }
  double Qb[6] = {0};
  for (int r = 0; r < rank; ++r)
    for (int m = 0; m < 6; ++m) Qb[r] += qr.Qt[r * 6 + m] * b[m];
  double G[36] = {0};  // R1 * R1^T (rank x rank)
  for (int r = 0; r < rank; ++r)
    for (int c = 0; c < rank; ++c) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s += R1[r * 6 + k] * R1[c * 6 + k];
      G[r * rank + c] = s;
    }
  double y[6] = {0};
  if (rank > 0) LltSolve(G, Qb, y, rank);
  double xp[6] = {0};  // R1.triangularView<Upper>().transpose() * y   (:230)
  for (int c = 0; c < 6; ++c)
    for (int r = 0; r < rank; ++r)
      if (c >= r) xp[c] += R1[r * 6 + c] * y[r];
  for (int c = 0; c < 6; ++c) x[c] = 0.0;
  for (int j = 0; j < 6; ++j) x[qr.perm[j]] = xp[j];  // x = P * xp   (:232)
  // :234-235  if (!b.isApprox(A x, 1e-5)) -> JacobiSVD
  double ax[6], nb = 0.0, nax = 0.0, ndiff = 0.0;
  for (int r = 0; r < 6; ++r) {
    ax[r] = 0.0;
    for (int c = 0; c < 6; ++c) ax[r] += A[r * 6 + c] * x[c];
    nb += b[r] * b[r]; nax += ax[r] * ax[r];
    ndiff += (b[r] - ax[r]) * (b[r] - ax[r]);
  }
  // isApprox: |b-ax|^2 <= prec^2 * min(|b|^2, |ax|^2)
  if (!(ndiff <= 1e-5 * 1e-5 * std::min(nb, nax))) {
    SymSvdSolve(A, b, x, 6);
    if (path) *path = 2;
  }
}

// icp_fast.cc:377-405.  quats: (w,x,y,z) per entry; trans: xyz per entry.
inline bool CheckConvergence(const std::vector<double>& quats,
                             const std::vector<double>& trans) {
  const size_t n = quats.size() / 4;
  constexpr size_t kSmoothLength = 4;
  if (n <= kSmoothLength) return false;
  double rot = 0.0, tr = 0.0;
  for (size_t i = n - 1; i >= n - kSmoothLength; --i) {
    rot += std::fabs(QuaternionAngularDistance(&quats[4 * i], &quats[4 * (i - 1)]));
    const double dx = trans[3 * i] - trans[3 * (i - 1)];
    const double dy = trans[3 * i + 1] - trans[3 * (i - 1) + 1];
    const double dz = trans[3 * i + 2] - trans[3 * (i - 1) + 2];
    tr += std::fabs(std::sqrt(dx * dx + dy * dy + dz * dz));
  }
  rot /= kSmoothLength;
  tr /= kSmoothLength;
  return rot < 0.001 && tr < 0.01;
}

}  // namespace
}  // namespace sm_oracle

namespace sm_oracle {
// Exact 1-NN of float points (converted to double; the k-d tree with eps = 0 is exact).
// Used for pcl::Registration::getFitnessScore (ndt.cc:60), whose FLANN search is exact.
int ExactNn1Float(const float* target_xyz, int64_t nt, const float* query_xyz, int64_t nq,
                  int32_t* ids_out) {
  std::vector<double> t((size_t)(3 * nt)), q((size_t)(3 * nq)), d2((size_t)nq);
  for (int64_t i = 0; i < 3 * nt; ++i) t[(size_t)i] = (double)target_xyz[i];
  for (int64_t i = 0; i < 3 * nq; ++i) q[(size_t)i] = (double)query_xyz[i];
  KdTree tree;
  tree.Build(t.data(), nt, 8, 0);
  tree.Knn1(q.data(), nq, 0.0, ids_out, d2.data());
  return 0;
}
// Exact k-NN of float points (pcl::search::KdTree::nearestKSearch, gicp_omp_impl.hpp:88).
int ExactKnnFloat(const float* target_xyz, int64_t nt, const float* query_xyz, int64_t nq, int k,
                  int32_t* ids_out, double* d2_out) {
  std::vector<double> t((size_t)(3 * nt)), q((size_t)(3 * nq));
  for (int64_t i = 0; i < 3 * nt; ++i) t[(size_t)i] = (double)target_xyz[i];
  for (int64_t i = 0; i < 3 * nq; ++i) q[(size_t)i] = (double)query_xyz[i];
  KdTree tree;
  tree.Build(t.data(), nt, 8, 0);
  tree.KnnK(q.data(), nq, k, ids_out, d2_out);
  return 0;
}
}  // namespace sm_oracle

using namespace sm_oracle;

extern "C" {

int64_t sm_oracle_calculate_normals(double* points_io, double* normals_out, int64_t n,
                                    int tie_mode) {
  if (n <= 0) return 0;
  NormalsBuilder nb;
  nb.points = points_io;
  nb.normals = normals_out;
  nb.tie_mode = tie_mode;
  nb.indices.resize((size_t)n);
  std::iota(nb.indices.begin(), nb.indices.end(), 0);
  double minv[3] = {kInf, kInf, kInf}, maxv[3] = {-kInf, -kInf, -kInf};
  for (int64_t i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r) {
      minv[r] = std::min(minv[r], points_io[3 * i + r]);
      maxv[r] = std::max(maxv[r], points_io[3 * i + r]);
    }
  nb.Build(0, (int)n, minv, maxv);
  // cloud_types.cc:357-367: sort survivors, compact to the front.
  std::sort(nb.indices_to_keep.begin(), nb.indices_to_keep.end());
  const int64_t m = (int64_t)nb.indices_to_keep.size();
  for (int64_t i = 0; i < m; ++i) {
    const int k = nb.indices_to_keep[(size_t)i];
    for (int r = 0; r < 3; ++r) {
      points_io[3 * i + r] = points_io[3 * (size_t)k + r];
      normals_out[3 * i + r] = normals_out[3 * (size_t)k + r];
    }
  }
  return m;
}

int sm_oracle_knn1(const double* target, int64_t n_target, const double* query,
                   int64_t n_query, double epsilon, int bucket_size, int tie_mode,
                   int32_t* ids_out, double* dists2_out) {
  if (n_target < 0 || n_query < 0 || bucket_size < 2) return -1;
  KdTree tree;
  tree.Build(target, n_target, bucket_size, tie_mode);
  tree.Knn1(query, n_query, epsilon, ids_out, dists2_out);
  return 0;
}

int sm_oracle_knn1_visits(const double* target, int64_t n_target, const double* query,
                          int64_t n_query, double epsilon, int bucket_size, int32_t* visits_out) {
  KdTree tree;
  tree.Build(target, n_target, bucket_size, 0);
  std::vector<int32_t> ids((size_t)n_query);
  std::vector<double> d2((size_t)n_query);
  tree.Knn1(query, n_query, epsilon, ids.data(), d2.data(), visits_out);
  return 0;
}

int sm_oracle_knn1_brute(const double* target, int64_t n_target, const double* query,
                         int64_t n_query, int32_t* ids_out, double* dists2_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_query; ++i) {
    double best = kInf;
    int best_idx = -1;
    for (int64_t j = 0; j < n_target; ++j) {
      double dist = 0.0;
      for (int r = 0; r < 3; ++r) {
        const double diff = query[3 * i + r] - target[3 * j + r];
        dist += diff * diff;
      }
      if (dist < best) { best = dist; best_idx = (int)j; }
    }
    ids_out[i] = best_idx;
    dists2_out[i] = best;
  }
  return 0;
}

int sm_oracle_icp_fast_align(const double* source, int64_t n_source, const double* target,
                             const double* target_normals, int64_t n_target,
                             const double* guess, const sm_oracle_icp_options* opt,
                             double* result, double* final_score, int32_t* iterations,
                             sm_oracle_icp_trace* trace, int32_t trace_capacity) {
  if (n_source <= 0 || n_target <= 0) return -1;
  const int64_t ns = n_source, nt = n_target;
  // :456-463 target mean (rowwise sum is a sequential loop for a strided row), centre.
  double mean[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < 3; ++r) {
    double s = 0.0;
    for (int64_t j = 0; j < nt; ++j) s += target[3 * j + r];
    mean[r] = s / (double)(int)nt;
  }
  std::vector<double> Q((size_t)(3 * nt));
  for (int64_t j = 0; j < nt; ++j)
    for (int r = 0; r < 3; ++r) Q[(size_t)(3 * j + r)] = target[3 * j + r] - mean[r];
  // :466-467 rebuild the tree on every Align.
  KdTree tree;
  tree.Build(Q.data(), nt, 8, opt->tie_mode);
  // :460-461,:469-471  T_mean, G0 = T_mean^-1 * guess, init_source = G0 (x) source.
  double T_mean[16], T_mean_inv[16], G0[16];
  Identity4(T_mean); Identity4(T_mean_inv);
  for (int r = 0; r < 3; ++r) { T_mean[12 + r] = mean[r]; T_mean_inv[12 + r] = -mean[r]; }
  Mul4(T_mean_inv, guess, G0);
  std::vector<double> S0((size_t)(3 * ns)), P((size_t)(3 * ns));
  ApplyTransform(G0, source, S0.data(), ns);

  double T_iter[16];
  Identity4(T_iter);
  std::vector<double> quats = {1.0, 0.0, 0.0, 0.0};
  std::vector<double> trans = {0.0, 0.0, 0.0};
  std::vector<int32_t> ids((size_t)ns);
  std::vector<double> d2((size_t)ns), values;
  values.reserve((size_t)ns);
  int iterator = 0;
  while (true) {
    ApplyTransform(T_iter, S0.data(), P.data(), ns);                      // :486-491
    tree.Knn1(P.data(), ns, opt->knn_epsilon, ids.data(), d2.data());     // :493
    // :65-90 quantile of the finite squared distances
    values.clear();
    for (int64_t i = 0; i < ns; ++i)
      if (d2[(size_t)i] != kInf) values.push_back(d2[(size_t)i]);
    if (values.empty()) return -2;
    const double quantile = (double)opt->dist_outlier_ratio;
    double limit;
    if (quantile == 1.0) {
      limit = *std::max_element(values.begin(), values.end());
    } else {
      const int qi = QuantileIndex((int64_t)values.size(), opt->dist_outlier_ratio);
      std::nth_element(values.begin(), values.begin() + qi, values.end());
      limit = values[(size_t)qi];
    }
    // :497-498 weights; :113-156 compaction + gathers; :256-302 normal equations.
    double A[36] = {0}, b[6] = {0};
    double sum_sqrt = 0.0;
    int64_t kept = 0;
    for (int64_t i = 0; i < ns; ++i) {
      const double dist = d2[(size_t)i];
      if (dist == kInf) continue;
      if (!(dist <= limit)) continue;
      const double* p = &P[(size_t)(3 * i)];
      const double* q = &Q[(size_t)(3 * (int64_t)ids[(size_t)i])];
      const double* nrm = target_normals + 3 * (int64_t)ids[(size_t)i];
      double F[6];
      F[0] = p[1] * nrm[2] - p[2] * nrm[1];   // :193-198
      F[1] = p[2] * nrm[0] - p[0] * nrm[2];
      F[2] = p[0] * nrm[1] - p[1] * nrm[0];
      F[3] = nrm[0]; F[4] = nrm[1]; F[5] = nrm[2];
      double dot = 0.0;                        // :296-299
      dot += (p[0] - q[0]) * nrm[0];
      dot += (p[1] - q[1]) * nrm[1];
      dot += (p[2] - q[2]) * nrm[2];
      for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 6; ++c) A[r * 6 + c] += F[r] * F[c];   // :292 (w == 1)
        b[r] += F[r] * dot;
      }
      sum_sqrt += std::sqrt(dist);
      ++kept;
    }
    if (kept == 0) return -3;
    for (int r = 0; r < 6; ++r) b[r] = -b[r];  // :302
    double x[6];
    SolvePossiblyUnderdetermined(A, b, x, nullptr);
    // :307-321 parameters -> 4x4
    const double angle = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    double axis[3] = {x[0], x[1], x[2]};
    const double sq = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    if (sq > 0.0) for (int r = 0; r < 3; ++r) axis[r] = x[r] / std::sqrt(sq);
    double R[9];
    AngleAxisToRotation(angle, axis, R);
    bool has_nan = false;
    for (int i = 0; i < 9; ++i) has_nan |= std::isnan(R[i]);
    for (int i = 3; i < 6; ++i) has_nan |= std::isnan(x[i]);
    if (has_nan) { for (int i = 0; i < 9; ++i) R[i] = 0.0; R[0] = R[4] = R[8] = 1.0; }
    double dT[16];
    Identity4(dT);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) dT[r + 4 * c] = R[r * 3 + c];
      dT[12 + r] = x[3 + r];
    }
    Mul4(dT, T_iter, T_iter);                   // :506-510
    if (trace && iterator < trace_capacity) {
      sm_oracle_icp_trace& t = trace[iterator];
      std::memcpy(t.T_iter, T_iter, sizeof(T_iter));
      t.limit = limit; t.kept = kept;
      std::memcpy(t.A, A, sizeof(A)); std::memcpy(t.b, b, sizeof(b));
    }
    ++iterator;                                 // :513-515
    double Rm[9], qn[4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rm[r * 3 + c] = T_iter[r + 4 * c];
    RotationToQuaternion(Rm, qn);
    quats.insert(quats.end(), qn, qn + 4);
    trans.push_back(T_iter[12]); trans.push_back(T_iter[13]); trans.push_back(T_iter[14]);
    const bool conv = !opt->disable_convergence_check && CheckConvergence(quats, trans);
    if (conv || iterator >= opt->max_iteration) {   // :516-522
      *final_score = std::exp(-(sum_sqrt / (double)kept));
      break;
    }
  }
  double tmp[16];
  // :527  Eigen evaluates T_mean * T_iter * G0 left to right: (T_mean*T_iter)*G0.
  Mul4(T_mean, T_iter, tmp);
  Mul4(tmp, G0, result);
  *iterations = iterator;
  return 1;
}

int sm_oracle_solve6(const double* A_colmajor, const double* b, double* x, int* path) {
  double A[36];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) A[r * 6 + c] = A_colmajor[r + 6 * c];
  SolvePossiblyUnderdetermined(A, b, x, path);
  return 0;
}

int sm_oracle_quantile_index(int64_t n, float ratio) { return QuantileIndex(n, ratio); }

void sm_oracle_check_convergence_inputs(const double* Ts, int n, int* converged) {
  std::vector<double> quats = {1.0, 0.0, 0.0, 0.0}, trans = {0.0, 0.0, 0.0};
  *converged = 0;
  for (int i = 0; i < n; ++i) {
    const double* T = Ts + 16 * i;
    double Rm[9], q[4];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rm[r * 3 + c] = T[r + 4 * c];
    RotationToQuaternion(Rm, q);
    quats.insert(quats.end(), q, q + 4);
    trans.push_back(T[12]); trans.push_back(T[13]); trans.push_back(T[14]);
    if (CheckConvergence(quats, trans)) { *converged = i + 1; return; }
  }
}

void sm_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int sm_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
