// Stand-ins for the reference headers the adapter needs (registrators/interface.h,
// builder/data/cloud_types.h, Eigen, glog), which are not available in this repository.  They are
// functional — small but real containers and an aborting CHECK — so that
// adapter/registrators_b200.h can be syntax-checked (tests/test_adapter_syntax.py) AND compiled
// into a test program that drives the C ABI through the C++ surface (tests/cpp/adapter_run.cc).
// Semantics copied from the reference: Interface::SetInputSource/Target keep the shared_ptr
// (interface.cc:38-60); CHECK aborts with file:line like glog.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
namespace Eigen {
// column-major dynamic matrix of doubles (only what the adapter touches)
struct MatrixXd {
  MatrixXd() = default;
  MatrixXd(long r, long c) : rows_(r), cols_(c), v_((size_t)(r * c), 0.0) {}
  const double* data() const { return v_.data(); }
  double* data() { return v_.data(); }
  long rows() const { return rows_; }
  long cols() const { return cols_; }
  double& operator()(long r, long c) { return v_[(size_t)(c * rows_ + r)]; }
  double operator()(long r, long c) const { return v_[(size_t)(c * rows_ + r)]; }
  MatrixXd leftCols(long m) const {
    MatrixXd o(rows_, m);
    for (long i = 0; i < rows_ * m; ++i) o.v_[(size_t)i] = v_[(size_t)i];
    return o;
  }
 private:
  long rows_ = 0, cols_ = 0;
  std::vector<double> v_;
};
struct Matrix4d {   // column-major, like Eigen's default
  Matrix4d() { for (double& x : d) x = 0.0; }
  static Matrix4d Identity() { Matrix4d m; m.d[0] = m.d[5] = m.d[10] = m.d[15] = 1.0; return m; }
  const double* data() const { return d; }
  double* data() { return d; }
  double& operator()(int r, int c) { return d[c * 4 + r]; }
  double operator()(int r, int c) const { return d[c * 4 + r]; }
  double d[16];
};
}  // namespace Eigen
// glog-like CHECK: streams a message and aborts when the condition is false
struct FatalStream {
  FatalStream(const char* file, int line, const char* what) { s_ << file << ":" << line << " Check failed: " << what << " "; }
  ~FatalStream() { std::cerr << s_.str() << std::endl; std::abort(); }
  template <typename T> FatalStream& operator<<(const T& v) { s_ << v; return *this; }
  std::ostringstream s_;
};
#define CHECK(x) if (!(x)) FatalStream(__FILE__, __LINE__, #x)
#define CHECK_EQ(a, b) if (!((a) == (b))) FatalStream(__FILE__, __LINE__, #a " == " #b)
#define CHECK_GE(a, b) if (!((a) >= (b))) FatalStream(__FILE__, __LINE__, #a " >= " #b)
#define CHECK_NE(a, b) if (!((a) != (b))) FatalStream(__FILE__, __LINE__, #a " != " #b)
struct LogStream {   // LOG(ERROR) of glog: prints, does not abort
  ~LogStream() { fprintf(stderr, "\n"); }
  template <typename T> LogStream& operator<<(const T& v) { std::cerr << v; return *this; }
};
#define ERROR 0
#define LOG(severity) LogStream()
#define PROHIBIT_COPY_AND_ASSIGN(C) C(const C&) = delete; C& operator=(const C&) = delete
namespace static_map {
namespace data {
struct EigenPointCloud {   // builder/data/cloud_types.h:121-147
  Eigen::MatrixXd points, normals;
  bool HasNormals() const { return normals.cols() > 0 && normals.cols() == points.cols(); }
};
struct InnerPointType { float x = 0, y = 0, z = 0, intensity = 0, factor = 0; };   // cloud_types.h:46-52
struct InnerCloudType { long long stamp = 0; std::vector<InnerPointType> points; };   // stamp: SimpleTime in the reference
struct InnerPointCloudData {
  using Ptr = std::shared_ptr<InnerPointCloudData>;
  InnerPointCloudData() : eigen_(new EigenPointCloud), inner_(new InnerCloudType) {}
  std::shared_ptr<EigenPointCloud> GetEigenCloud() const { return eigen_; }
  std::shared_ptr<InnerCloudType> GetInnerCloud() const { return inner_; }
  bool Empty() const { return inner_->points.empty(); }
 private:
  std::shared_ptr<EigenPointCloud> eigen_;
  std::shared_ptr<InnerCloudType> inner_;
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
  virtual void SetInputSource(InnerCloudPtr cloud) { source_cloud_ = cloud; }   // interface.cc:38-48
  virtual void SetInputTarget(InnerCloudPtr cloud) { target_cloud_ = cloud; }   // interface.cc:50-60
  virtual bool Align(const Eigen::Matrix4d&, Eigen::Matrix4d&) = 0;
  virtual double GetFitnessScore() { return final_score_; }
  virtual Type GetType() const { return type_; }
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
