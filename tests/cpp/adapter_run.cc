// Drives libsm_b200.so through the C++ surface a StaticMapping maintainer would use
// (adapter/registrators_b200.h on top of registrator::Interface), with the stand-in headers of
// tests/stubs.  Reads   <in>:  int64 ns, int64 nt, float src[ns*3], float tgt[nt*3], double delta[16]
// writes <out> (text): one "key v0 v1 ..." line per result.  tests/test_adapter_cpp.py builds it,
// runs it on the GPU and compares every line with the oracle.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <memory>
#include <vector>

#include "registrators_b200.h"

using namespace static_map;

static void Dump(std::ofstream& o, const char* key, const double* v, int n) {
  o << key;
  for (int i = 0; i < n; ++i) o << " " << std::setprecision(17) << v[i];
  o << "\n";
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int64_t ns = 0, nt = 0;
  in.read(reinterpret_cast<char*>(&ns), 8);
  in.read(reinterpret_cast<char*>(&nt), 8);
  std::vector<float> src(3 * ns), tgt(3 * nt);
  in.read(reinterpret_cast<char*>(src.data()), src.size() * 4);
  in.read(reinterpret_cast<char*>(tgt.data()), tgt.size() * 4);
  Eigen::Matrix4d delta;
  in.read(reinterpret_cast<char*>(delta.data()), 16 * 8);
  if (!in) return 3;
  std::ofstream out(argv[2]);

  // the clouds as the reference holds them: float AoS + double 3xN (EigenPointCloud::FromPointCloud)
  auto make_cloud = [](const std::vector<float>& xyz, int64_t n) {
    auto c = std::make_shared<data::InnerPointCloudData>();
    c->GetInnerCloud()->points.resize(n);
    c->GetEigenCloud()->points = Eigen::MatrixXd(3, n);
    for (int64_t i = 0; i < n; ++i) {
      auto& p = c->GetInnerCloud()->points[i];
      p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2];
      p.intensity = static_cast<float>(i % 251);
      p.factor = n > 1 ? static_cast<float>(i) / static_cast<float>(n - 1) : 0.f;
      for (int d = 0; d < 3; ++d) c->GetEigenCloud()->points(d, i) = static_cast<double>(xyz[3 * i + d]);
    }
    return c;
  };
  auto source = make_cloud(src, ns), target = make_cloud(tgt, nt);

  // target prep (map_builder.cc:286,389) then IcpFast through the Interface
  registrator::CalculateNormalsB200(target->GetEigenCloud().get());
  const double m = static_cast<double>(target->GetEigenCloud()->points.cols());
  Dump(out, "normals_count", &m, 1);
  {
    std::unique_ptr<registrator::Interface> matcher(new registrator::IcpFastB200());
    matcher->InitWithOptions();
    matcher->SetInputSource(source);
    matcher->SetInputTarget(target);
    Eigen::Matrix4d result;
    const bool ok = matcher->Align(Eigen::Matrix4d::Identity(), result);
    const double meta[3] = {ok ? 1.0 : 0.0, matcher->GetFitnessScore(), static_cast<double>(matcher->GetType())};
    Dump(out, "icp_result", result.data(), 16);
    Dump(out, "icp_meta", meta, 3);
  }
  {
    // batched CloseLoop sketch: three IcpFast candidates aligned by ONE thread through AlignBatch
    std::vector<std::unique_ptr<registrator::IcpFastB200>> owned;
    std::vector<registrator::IcpFastB200*> ms;
    std::vector<Eigen::Matrix4d> guesses, results;
    for (int k = 0; k < 3; ++k) {
      owned.emplace_back(new registrator::IcpFastB200());
      owned.back()->InitWithOptions();
      owned.back()->SetInputSource(source);
      owned.back()->SetInputTarget(target);
      ms.push_back(owned.back().get());
      Eigen::Matrix4d g = Eigen::Matrix4d::Identity();
      g(0, 3) = 0.01 * k;                              // distinct guesses
      guesses.push_back(g);
    }
    std::vector<bool> ok;
    registrator::AlignBatch(ms, guesses, &results, &ok);
    for (int k = 0; k < 3; ++k) {
      Eigen::Matrix4d single;
      std::unique_ptr<registrator::Interface> one(new registrator::IcpFastB200());
      one->InitWithOptions(); one->SetInputSource(source); one->SetInputTarget(target);
      const bool ok1 = one->Align(guesses[k], single);
      double diff = 0.0;
      for (int i = 0; i < 16; ++i) diff += std::fabs(single.data()[i] - results[k].data()[i]);
      const double rec[4] = {static_cast<double>(ok[k]), static_cast<double>(ok1), diff,
                             ms[k]->GetFitnessScore() - one->GetFitnessScore()};
      Dump(out, k == 0 ? "batch0" : (k == 1 ? "batch1" : "batch2"), rec, 4);
    }
  }
  {
    std::unique_ptr<registrator::Interface> matcher(new registrator::NdtB200());
    matcher->InitWithOptions();
    matcher->SetInputSource(source);
    matcher->SetInputTarget(target);
    Eigen::Matrix4d result;
    const bool ok = matcher->Align(Eigen::Matrix4d::Identity(), result);
    const double meta[3] = {ok ? 1.0 : 0.0, matcher->GetFitnessScore(), static_cast<double>(matcher->GetType())};
    Dump(out, "ndt_result", result.data(), 16);
    Dump(out, "ndt_meta", meta, 3);
  }
  {
    std::unique_ptr<registrator::Interface> matcher(new registrator::IcpUsingPointMatcherB200());
    matcher->SetInputSource(source);
    matcher->SetInputTarget(target);
    Eigen::Matrix4d result;
    const bool ok = matcher->Align(Eigen::Matrix4d::Identity(), result);
    const double meta[3] = {ok ? 1.0 : 0.0, matcher->GetFitnessScore(), static_cast<double>(matcher->GetType())};
    Dump(out, "pm_result", result.data(), 16);
    Dump(out, "pm_meta", meta, 3);
  }
  {
    data::InnerCloudType comp;
    registrator::MotionCompensationB200(*source->GetInnerCloud(), delta, &comp);
    double sum[5] = {0, 0, 0, 0, 0};
    for (const auto& p : comp.points) { sum[0] += p.x; sum[1] += p.y; sum[2] += p.z; sum[3] += p.intensity; sum[4] += p.factor; }
    Dump(out, "motion_sums", sum, 5);
    const auto& q = comp.points[comp.points.size() / 2];
    const double mid[3] = {q.x, q.y, q.z};
    Dump(out, "motion_mid", mid, 3);
  }
  {
    data::InnerCloudType vox;
    registrator::VoxelGridFilterB200(*target->GetInnerCloud(), 0.5f, &vox);
    double sum[5] = {static_cast<double>(vox.points.size()), 0, 0, 0, 0};
    for (const auto& p : vox.points) { sum[1] += p.x; sum[2] += p.y; sum[3] += p.z; sum[4] += p.intensity; }
    Dump(out, "voxel_sums", sum, 5);
  }
  return 0;
}
