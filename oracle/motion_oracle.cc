// ORACLE — TEST INFRASTRUCTURE ONLY (see sm_oracle.h).  PARITY UNPINNED: no reference test covers
// these functions; checked against analytic known answers (tests/test_oracle_motion.py).
//
// CPU restatement of the motion-compensation step either side of Align:
//   MotionCompensation           builder/map_builder.cc:232-257
//   common::InterpolateTransform common/math.h:198-211
//   common::AverageTransforms    common/math.cc:177-195 (RotationMatrixToEulerAngles
//                                common/math.h:107-127, EulerAnglesToQuaternion :129-139)
// and of the Eigen 3.3 pieces they call (Quaternion(Matrix3), QuaternionBase::slerp,
// QuaternionBase::toRotationMatrix, AngleAxis -> Quaternion, quaternion product).
// All 4x4 matrices are column-major doubles.
#include <cmath>
#include <cstdint>

#include "linalg.h"
#include "sm_oracle.h"

namespace sm_oracle {
namespace {

struct Quat { double w, x, y, z; };

Quat QuatFromTransform(const double* T) {
  double m[9], q[4];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[r * 3 + c] = T[r + 4 * c];
  RotationToQuaternion(m, q);
  return Quat{q[0], q[1], q[2], q[3]};
}

// QuaternionBase::slerp (Eigen/src/Geometry/Quaternion.h)
Quat Slerp(const Quat& a, double t, const Quat& b) {
  const double one = 1.0 - 2.220446049250313e-16;
  const double d = ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w;
  const double abs_d = std::fabs(d);
  double scale0, scale1;
  if (abs_d >= one) {
    scale0 = 1.0 - t;
    scale1 = t;
  } else {
    const double theta = std::acos(abs_d);
    const double sin_theta = std::sin(theta);
    scale0 = std::sin((1.0 - t) * theta) / sin_theta;
    scale1 = std::sin(t * theta) / sin_theta;
  }
  if (d < 0.0) scale1 = -scale1;
  return Quat{scale0 * a.w + scale1 * b.w, scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y,
              scale0 * a.z + scale1 * b.z};
}

// QuaternionBase::toRotationMatrix; R row-major
void QuatToRotation(const Quat& q, double* R) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

Quat QuatMul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

// common/math.h:198-211
void InterpolateTransform(const double* t1, const double* t2, float factor, double* out) {
  Identity4(out);
  const Quat qa = QuatFromTransform(t1), qb = QuatFromTransform(t2);
  double R[9];
  QuatToRotation(Slerp(qa, (double)factor, qb), R);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out[r + 4 * c] = R[r * 3 + c];
  for (int r = 0; r < 3; ++r) out[12 + r] = t1[12 + r] + (t2[12 + r] - t1[12 + r]) * (double)factor;
}

// common/math.h:107-127
void RotationMatrixToEulerAngles(const double* T, double* e) {
  auto R = [&](int r, int c) { return T[r + 4 * c]; };
  const double sy = std::sqrt(R(0, 0) * R(0, 0) + R(1, 0) * R(1, 0));
  if (!(sy < 1e-6)) {
    e[0] = std::atan2(R(2, 1), R(2, 2));
    e[1] = std::atan2(-R(2, 0), sy);
    e[2] = std::atan2(R(1, 0), R(0, 0));
  } else {
    e[0] = std::atan2(-R(1, 2), R(1, 1));
    e[1] = std::atan2(-R(2, 0), sy);
    e[2] = 0.0;
  }
}

Quat QuatFromAngleAxis(double angle, int axis) {
  const double h = 0.5 * angle, s = std::sin(h);
  Quat q{std::cos(h), 0.0, 0.0, 0.0};
  (axis == 0 ? q.x : axis == 1 ? q.y : q.z) = s * 1.0;
  return q;
}

}  // namespace
}  // namespace sm_oracle

using namespace sm_oracle;

extern "C" {

int sm_oracle_interpolate_transform(const double* t1, const double* t2, float factor, double* out) {
  if (!(factor >= 0.f && factor <= 1.f)) return -1;   // CHECK, common/math.h:201
  InterpolateTransform(t1, t2, factor, out);
  return 0;
}

// points / out: packed InnerPointType (x, y, z, intensity, factor), 5 floats per point
int sm_oracle_motion_compensation(const float* points, int64_t n, const double* delta, float* out) {
  double ident[16], T[16];
  Identity4(ident);
  for (int64_t i = 0; i < n; ++i) {
    const float* p = points + 5 * i;
    if (!(p[4] >= 0.f && p[4] <= 1.f)) return -1;
    InterpolateTransform(ident, delta, p[4], T);
    const double x = p[0], y = p[1], z = p[2];
    float* o = out + 5 * i;
    for (int r = 0; r < 3; ++r)
      o[r] = (float)(((T[r] * x + T[r + 4] * y) + T[r + 8] * z) + T[12 + r]);
    o[3] = p[3];
    o[4] = p[4];
  }
  return 0;
}

// common/math.cc:177-195
int sm_oracle_average_transforms(const double* Ts, int32_t n, double* out) {
  if (n <= 0) return -1;
  double ang[3] = {0, 0, 0}, tr[3] = {0, 0, 0};
  for (int k = 0; k < n; ++k) {
    const double* T = Ts + 16 * k;
    double e[3];
    RotationMatrixToEulerAngles(T, e);
    for (int d = 0; d < 3; ++d) { tr[d] += T[12 + d]; ang[d] += e[d]; }
  }
  for (int d = 0; d < 3; ++d) { tr[d] /= (double)n; ang[d] /= (double)n; }
  // yaw * pitch * roll (common/math.h:129-139)
  const Quat q = QuatMul(QuatMul(QuatFromAngleAxis(ang[2], 2), QuatFromAngleAxis(ang[1], 1)),
                         QuatFromAngleAxis(ang[0], 0));
  double R[9];
  QuatToRotation(q, R);
  Identity4(out);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) out[r + 4 * c] = R[r * 3 + c];
    out[12 + r] = tr[r];
  }
  return 0;
}

}  // extern "C"

// ---- pre_processers::filter::VoxelGrid::Filter (pre_processors/filter_voxel_grid.cc:37-78) --------
// PINNED by the reference's own unit test (pre_processors/test/test_filter_voxel_grid.cc:34-100: 100 / 36 / 9
// voxels for its 10 x 10 lattice at 0.1 / 0.2 / 0.4, voxel_size 0 invalid), tests/test_oracle_voxel_filter.py.
// points / out: packed InnerPointType (x, y, z, intensity, factor).  order_mode 0: output voxels in
// ascending (ix, iy, iz) (the order the CUDA path emits); order_mode 1: the literal iteration order
// of std::unordered_map<Eigen::Vector3i, ...> with the reference's hash (common/eigen_hash.h:31-43)
// as this toolchain's libstdc++ produces it.  Returns the number of voxels, -1 on a bad voxel size.
#include <array>
#include <cmath>
#include <map>
#include <unordered_map>
#include <vector>

namespace {
struct Key3 { int v[3]; bool operator==(const Key3& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; } };
struct Key3Hash {   // common/eigen_hash.h:34-42 with std::hash<int> (identity)
  size_t operator()(const Key3& k) const {
    size_t seed = 0;
    for (int i = 0; i < 3; ++i) seed ^= std::hash<int>()(k.v[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};
void AveragePoint(const std::vector<const float*>& pts, float* o) {   // filter_voxel_grid.cc:54-72
  double sum[4] = {0, 0, 0, 0};
  for (const float* p : pts) { sum[0] += p[0]; sum[1] += p[1]; sum[2] += p[2]; sum[3] += p[3]; }
  const int size = (int)pts.size();
  o[0] = (float)(sum[0] / size); o[1] = (float)(sum[1] / size); o[2] = (float)(sum[2] / size);
  o[3] = (float)(sum[3] / size);
  o[4] = 0.f;   // InnerPointType::factor default
}
}  // namespace

extern "C" int64_t sm_oracle_voxel_grid_filter(const float* points, int64_t n, float voxel_size, int order_mode,
                                                float* out) {
  if (!(voxel_size > 1.e-6f)) return -1;   // ConfigsValid(), filter_voxel_grid.cc:35
  int64_t m = 0;
  auto key_of = [&](const float* p) {
    return Key3{{(int)std::lround(p[0] / voxel_size), (int)std::lround(p[1] / voxel_size),
                 (int)std::lround(p[2] / voxel_size)}};
  };
  if (order_mode == 1) {
    std::unordered_map<Key3, std::vector<const float*>, Key3Hash> grid;
    for (int64_t i = 0; i < n; ++i) grid[key_of(points + 5 * i)].push_back(points + 5 * i);
    for (const auto& g : grid) AveragePoint(g.second, out + 5 * (m++));
  } else {
    std::map<std::array<int, 3>, std::vector<const float*>> grid;
    for (int64_t i = 0; i < n; ++i) {
      const Key3 k = key_of(points + 5 * i);
      grid[{k.v[0], k.v[1], k.v[2]}].push_back(points + 5 * i);
    }
    for (const auto& g : grid) AveragePoint(g.second, out + 5 * (m++));
  }
  return m;
}
