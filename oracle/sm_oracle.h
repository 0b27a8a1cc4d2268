/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (double precision, reference operation order) of StaticMapping's
 * registrators/ hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (staticmapping_b200/, libsm_b200.so) never does.
 *
 * PARITY UNPINNED: the reference ships no test, golden vector or fixture for
 * registrators/ (README.md:203-206; builder/data/test/test_cloud_types.cc:158 is an
 * empty stub) and cannot be compiled in this image (Eigen, PCL, libnabo, glog, Boost
 * absent), so this restatement is checked only against analytic known-answer scenes,
 * brute force and scipy's cKDTree (tests/test_oracle_*.py) and against a second, independent
 * restatement of the same reference algorithms in Python / numpy (tests/pyref.py,
 * tests/test_oracle_vs_python_restatement.py).  The one row that does have reference
 * golden values is the voxel filter (pre_processors/test/test_filter_voxel_grid.cc), which the
 * restatement reproduces (tests/test_oracle_voxel_filter.py).
 *
 * All matrices crossing this API use Eigen's default layout: column-major.
 * Clouds are 3xN column-major doubles (x0,y0,z0,x1,...), as Eigen::MatrixXd in
 * data::EigenPointCloud (builder/data/cloud_types.h:143-146).
 */
#ifndef ORACLE_SM_ORACLE_H_
#define ORACLE_SM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* EigenPointCloud::CalculateNormals (builder/data/cloud_types.cc:347-368 with
 * BuildNormals :105-144 and the leaf routine :73-103).  points_io: 3xN in, 3xM out
 * (compacted in place like the reference); normals_out: 3xN capacity, 3xM written.
 * tie_mode 0: total-order comparator (coord, index) and leaf representative =
 * smallest original index (deterministic); tie_mode 1: literal std::nth_element with
 * the reference's coordinate-only comparator and representative = indices[first].
 * Returns M (number of surviving points). */
int64_t sm_oracle_calculate_normals(double* points_io, double* normals_out, int64_t n,
                                    int tie_mode);

/* libnabo 1.0.7 KDTREE_LINEAR_HEAP restatement (external dependency pinned by
 * setup/install_libnabo.sh:18; call sites registrators/icp_fast.cc:466-467, :177-178).
 * k = 1, ALLOW_SELF_MATCH, maxRadius = inf.  dists2_out are SQUARED distances;
 * ids_out = -1 / dists2 = +inf when nothing was found (NNS::InvalidIndex/Value).
 * bucket_size: 8 for libnabo's default.  Returns 0 on success. */
int sm_oracle_knn1(const double* target, int64_t n_target, const double* query,
                   int64_t n_query, double epsilon, int bucket_size, int tie_mode,
                   int32_t* ids_out, double* dists2_out);

/* Diagnostics: number of leaf buckets each query visits (traversal cost model). */
int sm_oracle_knn1_visits(const double* target, int64_t n_target, const double* query,
                          int64_t n_query, double epsilon, int bucket_size, int32_t* visits_out);

/* Exhaustive 1-NN (first minimum wins) — independent check for the two above. */
int sm_oracle_knn1_brute(const double* target, int64_t n_target, const double* query,
                         int64_t n_query, int32_t* ids_out, double* dists2_out);

typedef struct sm_oracle_icp_options {
  int32_t max_iteration;      /* icp_fast.h:58, default 100 */
  float dist_outlier_ratio;   /* icp_fast.h:59, default 0.7f (a FLOAT on purpose) */
  double knn_epsilon;         /* icp_fast.cc:174, 3.16 */
  int32_t disable_convergence_check; /* throughput runs: fixed iteration count */
  int32_t tie_mode;
} sm_oracle_icp_options;

typedef struct sm_oracle_icp_trace {   /* optional per-iteration record */
  double T_iter[16];
  double limit;
  int64_t kept;
  double A[36];
  double b[6];
} sm_oracle_icp_trace;

/* IcpFast::Align (registrators/icp_fast.cc:455-529).  source: 3xNs, target 3xNt with
 * unit normals 3xNt (the caller ran CalculateNormals, map_builder.cc:286,389).
 * Returns 1 (Align always returns true) or a negative code where the reference would
 * CHECK-fail: -1 empty input, -2 no finite match distance, -3 no point to minimise. */
int sm_oracle_icp_fast_align(const double* source, int64_t n_source, const double* target,
                             const double* target_normals, int64_t n_target,
                             const double* guess, const sm_oracle_icp_options* opt,
                             double* result, double* final_score, int32_t* iterations,
                             sm_oracle_icp_trace* trace, int32_t trace_capacity);

/* ---- NDT (registrators/ndt.cc + registrators/pclomp/, see oracle/ndt_oracle.cc) ---------- */
typedef struct sm_oracle_ndt_options {
  float resolution;              /* ndt.cc:31, 1.0 */
  double step_size;              /* ndt_omp_impl.hpp:49, 0.1 */
  double outlier_ratio;          /* :50, 0.55 */
  double transformation_epsilon; /* :71, 0.1 */
  int32_t max_iterations;        /* :72, 35 */
} sm_oracle_ndt_options;

/* Clouds here are packed float xyz (n x 3), i.e. pcl::PointXYZ without padding. */
int64_t sm_oracle_ndt_voxels(const float* target, int64_t n, float resolution, int64_t capacity,
                             int32_t* idx_out, int32_t* npts_out, double* mean_out,
                             double* icov_out, float* centroid_out, int32_t* searchable_out);
int sm_oracle_ndt_derivatives(const float* source, int64_t ns, const float* target, int64_t nt,
                              const sm_oracle_ndt_options* opt, const double* p, double* score,
                              double* grad6, double* hess36, double* mean_neighbors);
/* Ndt::Align (ndt.cc:38-64); returns 1 / 0 like the reference's bool. fitness = PCL
 * getFitnessScore (mean squared NN distance, lower is better). */
int sm_oracle_ndt_align(const float* source, int64_t ns, const float* target, int64_t nt,
                        const double* guess, const sm_oracle_ndt_options* opt, double* result,
                        double* fitness, int32_t* iterations, double* trans_probability,
                        double* mean_neighbors);

/* ---- NdtWithGicp (registrators/ndt_gicp.cc:28-112, see oracle/gicp_oracle.cc) ------------ */
typedef struct sm_oracle_ndt_gicp_options {
  float voxel_resolution;        /* ndt_gicp.h:72, 0.2 */
  int32_t using_voxel_filter;    /* :73, true */
  int32_t use_ndt;               /* :74, true */
} sm_oracle_ndt_gicp_options;
typedef struct sm_oracle_ndt_gicp_info {
  int64_t n_source_filtered, n_target_filtered;
  int32_t ndt_iterations, gicp_iterations, bfgs_evaluations, pad;
  double ndt_score, gicp_fitness;
} sm_oracle_ndt_gicp_info;
int64_t sm_oracle_approx_voxel_grid(const float* pts, int64_t n, float leaf, float* out, int64_t capacity);
int sm_oracle_gicp_covariances(const float* pts, int64_t n, int k, double eps, double* cov_out);
/* test hook: GICP's correspondence step + one cost / gradient evaluation (gicp_omp_impl.hpp:419-463,255-377) */
int64_t sm_oracle_gicp_cost(const float* src, int64_t ns, const float* tgt, int64_t nt, const float* guess,
                            const float* transformation, const double* x, double* f, double* g6, double* maha_out,
                            int32_t* si_out, int32_t* ti_out);
int sm_oracle_ndt_gicp_align(const float* source, int64_t ns, const float* target, int64_t nt,
                             const double* guess, const sm_oracle_ndt_gicp_options* opt, double* result,
                             double* final_score, sm_oracle_ndt_gicp_info* info);

/* ---- motion compensation either side of Align (oracle/motion_oracle.cc) -------------------- */
/* common::InterpolateTransform (common/math.h:198-211); -1 where the reference CHECK-fails. */
int sm_oracle_interpolate_transform(const double* t1, const double* t2, float factor, double* out);
/* MotionCompensation (builder/map_builder.cc:232-257); points/out: packed InnerPointType
 * (x, y, z, intensity, factor), 5 floats per point. */
int sm_oracle_motion_compensation(const float* points, int64_t n, const double* delta, float* out);
/* common::AverageTransforms (common/math.cc:177-195). */
int sm_oracle_average_transforms(const double* Ts, int32_t n, double* out);

/* pre_processers::filter::VoxelGrid::Filter (pre_processors/filter_voxel_grid.cc:37-78); packed
 * InnerPointType in/out; order_mode 0 = ascending (ix, iy, iz), 1 = the reference's literal
 * std::unordered_map iteration order.  Returns the number of voxels. */
int64_t sm_oracle_voxel_grid_filter(const float* points, int64_t n, float voxel_size, int order_mode,
                                    float* out);

/* ---- M2DP descriptor (descriptor/m2dp.cc:37-172, see oracle/m2dp_oracle.cc) ------------------ */
/* descriptor length p*q + l*t with l = ceil(sqrt(max_distance / r)) (m2dp.cc:68); -1 if r < 1e-6. */
int64_t sm_oracle_m2dp_dims(double r, double max_distance, int32_t t, int32_t p, int32_t q, int32_t* l_out);
/* M2dp::setInputCloud on packed float xyz; A_out (p*q x l*t counts) and axes_out (mean 3 + axes 9) optional. */
int64_t sm_oracle_m2dp(const float* points, int64_t n, double r, double max_distance, int32_t t, int32_t p,
                       int32_t q, float* descriptor, int32_t* A_out, float* axes_out);
/* matchTwoM2dpDescriptors (m2dp.cc:155-170). */
double sm_oracle_m2dp_match(const float* P, const float* Q, int64_t n);

/* Pieces exposed for unit tests of the restatement itself. */
int sm_oracle_solve6(const double* A_colmajor, const double* b, double* x, int* path);
int sm_oracle_quantile_index(int64_t n, float ratio);
void sm_oracle_check_convergence_inputs(const double* Ts, int n, int* converged);

int sm_oracle_num_threads(void);
/* OpenMP threads of the following calls (the NDT pieces keep the reference's own 6, ndt.cc:32). */
void sm_oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_SM_ORACLE_H_ */
