// Minimal stand-ins for the reference headers, ONLY to syntax-check adapter/registrators_b200.h
// in this repository (Eigen / glog / the reference tree are not available here).
#pragma once
#include <cstdint>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
namespace Eigen {
struct MatrixXd {
  MatrixXd() = default;
  MatrixXd(long, long) {}
  const double* data() const { return nullptr; }
  double* data() { return nullptr; }
  long cols() const { return 0; }
  MatrixXd leftCols(long) const { return *this; }
};
struct Matrix4d {
  const double* data() const { return d; }
  double* data() { return d; }
  double d[16];
};
}  // namespace Eigen
struct NullStream { template <typename T> NullStream& operator<<(const T&) { return *this; } };
#define CHECK(x) if (!(x)) NullStream()
#define CHECK_EQ(a, b) if (!((a) == (b))) NullStream()
#define CHECK_GE(a, b) if (!((a) >= (b))) NullStream()
#define PROHIBIT_COPY_AND_ASSIGN(C) C(const C&) = delete; C& operator=(const C&) = delete
namespace static_map {
namespace data {
struct EigenPointCloud {
  Eigen::MatrixXd points, normals;
  bool HasNormals() const { return true; }
};
struct InnerPointType { float x, y, z, intensity, factor; };
struct InnerCloudType { long long stamp = 0; std::vector<InnerPointType> points; };   // stamp: SimpleTime in the reference
struct InnerPointCloudData {
  using Ptr = std::shared_ptr<InnerPointCloudData>;
  std::shared_ptr<EigenPointCloud> GetEigenCloud() const { return nullptr; }
  std::shared_ptr<InnerCloudType> GetInnerCloud() const { return nullptr; }
};
}  // namespace data
namespace registrator {
enum Type { kNoType, kIcpPM, kLibicp, kNdtWithGicp, kLegoLoam, kNdt, kFastIcp, kTypeCount };
enum class OptionItemDataType : uint8_t { kInt32, kFloat32, kBool };
struct InnerOptionItem { OptionItemDataType data_type; void* data_ptr = nullptr; };
class Interface {
 public:
  using InnerCloudPtr = data::InnerPointCloudData::Ptr;
  Interface() = default;
  virtual ~Interface() {}
  virtual void InitWithOptions() {}
  virtual void SetInputSource(InnerCloudPtr) {}
  virtual void SetInputTarget(InnerCloudPtr) {}
  virtual bool Align(const Eigen::Matrix4d&, Eigen::Matrix4d&) = 0;
 protected:
  double final_score_ = 0;
  Type type_ = kNoType;
  InnerCloudPtr source_cloud_ = nullptr;
  InnerCloudPtr target_cloud_ = nullptr;
  std::unordered_map<std::string, InnerOptionItem> inner_options_;
};
}  // namespace registrator
}  // namespace static_map
#define USE_REGISTRATOR_CLOUDS using typename Interface::InnerCloudPtr;
#define REG_REGISTRATOR_INNER_OPTION(NAME, TYPE, VARIABLE) \
  this->inner_options_[NAME].data_type = TYPE;             \
  this->inner_options_[NAME].data_ptr = &VARIABLE;
