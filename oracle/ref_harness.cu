// ref_harness.cu — drives the UNMODIFIED reference PatchMatchCuda (compiled from /root/reference where it lies,
// against the stub headers in ref_stubs/) through the same flat structs as the product's C-ABI.
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (oracle/_ref): statistical cross-check of the oracle and the
// reference-arm timing of bench.py.  The functions below are the pieces of mvs/image.cc, depth_map.cc and
// normal_map.cc the CUDA code links against; they are re-implemented here without Eigen / OpenImageIO.
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "colmap/mvs/patch_match_cuda.h"
#include "../include/b200_patch_match.h"

namespace colmap {
namespace mvs {

Image::Image() {}
Image::Image(const std::filesystem::path& path, size_t width, size_t height, const float* K, const float* R, const float* T)
    : path_(path), width_(width), height_(height) {
  memcpy(K_, K, 9 * sizeof(float)); memcpy(R_, R, 9 * sizeof(float)); memcpy(T_, T, 3 * sizeof(float));
  ComposeProjectionMatrix(K_, R_, T_, P_);
  ComposeInverseProjectionMatrix(K_, R_, T_, inv_P_);
}
void Image::SetBitmap(Bitmap bitmap) { bitmap_ = std::move(bitmap); }

void ComputeRelativePose(const float R1[9], const float T1[3], const float R2[9], const float T2[3], float R[9], float T[3]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = R2[3 * r] * R1[3 * c] + R2[3 * r + 1] * R1[3 * c + 1] + R2[3 * r + 2] * R1[3 * c + 2];
  for (int r = 0; r < 3; ++r) T[r] = T2[r] - (R[3 * r] * T1[0] + R[3 * r + 1] * T1[1] + R[3 * r + 2] * T1[2]);
}
void ComposeProjectionMatrix(const float K[9], const float R[9], const float T[3], float P[12]) {
  float RT[12];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) RT[4 * r + c] = R[3 * r + c]; RT[4 * r + 3] = T[r]; }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) P[4 * r + c] = K[3 * r] * RT[c] + K[3 * r + 1] * RT[4 + c] + K[3 * r + 2] * RT[8 + c];
}
void ComposeInverseProjectionMatrix(const float K[9], const float R[9], const float T[3], float inv_P[12]) {
  // top three rows of [K[R|T]; 0 0 0 1]^-1 = [R^T K^-1 | -R^T T] (zero-skew K, as PatchMatch::Check enforces)
  const double fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  for (int r = 0; r < 3; ++r) {
    inv_P[4 * r + 0] = (float)(R[r] / fx);
    inv_P[4 * r + 1] = (float)(R[3 + r] / fy);
    inv_P[4 * r + 2] = (float)(R[6 + r] - R[r] * cx / fx - R[3 + r] * cy / fy);
    inv_P[4 * r + 3] = (float)(-((double)R[r] * T[0] + (double)R[3 + r] * T[1] + (double)R[6 + r] * T[2]));
  }
}
void ComputeProjectionCenter(const float R[9], const float T[3], float C[3]) {
  for (int c = 0; c < 3; ++c) C[c] = -(R[c] * T[0] + R[3 + c] * T[1] + R[6 + c] * T[2]);
}
void RotatePose(const float RR[9], float R[9], float T[3]) {
  float nR[9], nT[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) nR[3 * r + c] = RR[3 * r] * R[c] + RR[3 * r + 1] * R[3 + c] + RR[3 * r + 2] * R[6 + c];
    nT[r] = RR[3 * r] * T[0] + RR[3 * r + 1] * T[1] + RR[3 * r + 2] * T[2];
  }
  memcpy(R, nR, sizeof(nR)); memcpy(T, nT, sizeof(nT));
}

DepthMap::DepthMap() : DepthMap(0, 0, -1.0f, -1.0f) {}
DepthMap::DepthMap(size_t width, size_t height, float depth_min, float depth_max)
    : Mat<float>(width, height, 1), depth_min_(depth_min), depth_max_(depth_max) {}
DepthMap::DepthMap(const Mat<float>& mat, float depth_min, float depth_max)
    : Mat<float>(mat.GetWidth(), mat.GetHeight(), mat.GetDepth()), depth_min_(depth_min), depth_max_(depth_max) {
  data_ = mat.GetData();
}
NormalMap::NormalMap() : Mat<float>(0, 0, 3) {}
NormalMap::NormalMap(size_t width, size_t height) : Mat<float>(width, height, 3) {}
NormalMap::NormalMap(const Mat<float>& mat) : Mat<float>(mat.GetWidth(), mat.GetHeight(), mat.GetDepth()) { data_ = mat.GetData(); }

}  // namespace mvs
}  // namespace colmap

using namespace colmap;
using namespace colmap::mvs;

extern "C" int pm_ref_run(const b200pm_options* o, const b200pm_problem* p, float* depth, float* normal, float* sel_prob,
                          double* create_run_get_ms, char* err, int errlen) {
  try {
    PatchMatchOptions opt;
    opt.depth_min = o->depth_min; opt.depth_max = o->depth_max; opt.sigma_spatial = o->sigma_spatial;
    opt.sigma_color = o->sigma_color; opt.ncc_sigma = o->ncc_sigma; opt.min_triangulation_angle = o->min_triangulation_angle;
    opt.incident_angle_sigma = o->incident_angle_sigma; opt.geom_consistency_regularizer = o->geom_consistency_regularizer;
    opt.geom_consistency_max_cost = o->geom_consistency_max_cost; opt.filter_min_ncc = o->filter_min_ncc;
    opt.filter_min_triangulation_angle = o->filter_min_triangulation_angle;
    opt.filter_geom_consistency_max_cost = o->filter_geom_consistency_max_cost; opt.window_radius = o->window_radius;
    opt.window_step = o->window_step; opt.num_samples = o->num_samples; opt.num_iterations = o->num_iterations;
    opt.filter_min_num_consistent = o->filter_min_num_consistent; opt.geom_consistency = o->geom_consistency != 0;
    opt.filter = o->filter != 0; opt.gpu_index = std::to_string(o->gpu_index < 0 ? 0 : o->gpu_index);
    if (opt.sigma_spatial <= 0) opt.sigma_spatial = opt.window_radius;

    std::vector<Image> images;
    std::vector<DepthMap> depth_maps;
    std::vector<NormalMap> normal_maps;
    auto add = [&](int w, int h, const uint8_t* gray, const float* K, const float* R, const float* T) {
      images.emplace_back("", (size_t)w, (size_t)h, K, R, T);
      Bitmap bmp(w, h, false);
      memcpy(bmp.RowMajorData().data(), gray, (size_t)w * h);
      images.back().SetBitmap(std::move(bmp));
    };
    add(p->ref_width, p->ref_height, p->ref_gray, p->ref_K, p->ref_R, p->ref_T);
    for (int i = 0; i < p->num_src; ++i) add(p->src_width[i], p->src_height[i], p->src_gray[i], p->src_K + 9 * i, p->src_R + 9 * i, p->src_T + 3 * i);
    PatchMatch::Problem problem;
    problem.ref_image_idx = 0;
    for (int i = 0; i < p->num_src; ++i) problem.src_image_idxs.push_back(i + 1);
    problem.images = &images;
    if (opt.geom_consistency) {
      auto mk = [&](int w, int h, const float* d) { Mat<float> m(w, h, 1); memcpy(m.GetPtr(), d, sizeof(float) * w * h); return DepthMap(m, (float)opt.depth_min, (float)opt.depth_max); };
      depth_maps.push_back(mk(p->ref_width, p->ref_height, p->ref_depth_init));
      { Mat<float> m(p->ref_width, p->ref_height, 3); memcpy(m.GetPtr(), p->ref_normal_init, sizeof(float) * 3 * p->ref_width * p->ref_height); normal_maps.emplace_back(m); }
      for (int i = 0; i < p->num_src; ++i) { depth_maps.push_back(mk(p->src_width[i], p->src_height[i], p->src_depth[i])); normal_maps.emplace_back(p->src_width[i], p->src_height[i]); }
      problem.depth_maps = &depth_maps; problem.normal_maps = &normal_maps;
    }
    const auto t0 = std::chrono::steady_clock::now();
    PatchMatchCuda pm(opt, problem);
    pm.Run();
    const DepthMap d = pm.GetDepthMap();
    const NormalMap n = pm.GetNormalMap();
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    if (create_run_get_ms) *create_run_get_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    const size_t npx = (size_t)p->ref_width * p->ref_height;
    if (depth) memcpy(depth, d.GetPtr(), sizeof(float) * npx);
    if (normal) memcpy(normal, n.GetPtr(), sizeof(float) * 3 * npx);
    if (sel_prob) { const Mat<float> s = pm.GetSelProbMap(); memcpy(sel_prob, s.GetPtr(), sizeof(float) * npx * p->num_src); }
    return 0;
  } catch (const std::exception& e) {
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
}
