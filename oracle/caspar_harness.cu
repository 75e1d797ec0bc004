// caspar_harness.cu — flat-array driver of the REFERENCE's own GPU bundle-adjustment backend (Caspar).
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY (oracle/): nothing in the product may call, link or import this file.
//
// What it is: COLMAP's CasparBundleAdjuster (src/colmap/estimators/bundle_adjustment_caspar.cc) is a glue layer
// between `Reconstruction` and the generated solver src/thirdparty/Symforce-Caspar/generated/f32/solver.{h,cc} +
// 240 kernel files.  The glue needs Eigen / glog / Reconstruction, which this image lacks; the generated solver
// needs only CUDA.  oracle/build_caspar.sh compiles the generated sources from where they lie (never copied) and
// this file feeds them the stacked host arrays the glue would (SetupSolverData, :598-672; solve + read-back
// :926-970), from a b200ba_problem.  Same semantics as the glue: one pose pool per camera model, intrinsics split
// into focal_and_extra {f, k} / {fx, fy} and principal point, factor variants by which blocks are variable
// (AddFactorCore :285-376), the two-cams gauge fixes ONE frame only (FixGaugeWithOneFrameFromWorld :527-562: the
// second frame's translation coordinate stays free), fp32 everywhere, CasparBundleAdjustmentOptions defaults
// (bundle_adjustment_caspar.h:108-123).  Supported here: SIMPLE_RADIAL and PINHOLE (the two models Caspar has),
// refine_focal_length == refine_extra_params == true (Caspar's merged block), trivial frames.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../include/b200_bundle_adjustment.h"
#include "thirdparty/Symforce-Caspar/generated/f32/solver.h"

namespace {

struct Factors {   // one factor variant of one camera model (VariantData of the glue)
  std::vector<unsigned int> pose_idx, calib_idx, point_idx;
  std::vector<float> rig, const_pose, const_pp, const_point, pixels;
  size_t n = 0;
};
// variant key: bit2 = pose variable, bit1 = principal point variable, bit0 = point variable (focal_and_extra is
// always variable here)
struct ModelPool {
  std::vector<int> pose_of_image;        // image -> pose node or -1
  std::vector<float> pose_data;          // 7 per node
  std::vector<int> node_image;           // pose node -> image
  std::vector<int> calib_of_camera;      // camera -> calib node or -1
  std::vector<int> node_camera;
  std::vector<float> fae, pp;            // 2 + 2 per calib node
  Factors f[8];
};

#define SET_RIG(P) s.Set##P##SensorFromRigDataFromStackedHost(F.rig.data(), 0, n)
#define SET_PIX(P) s.Set##P##PixelDataFromStackedHost(F.pixels.data(), 0, n)
#define VAR_POSE(P) s.Set##P##PoseIndicesFromHost(F.pose_idx.data(), n)
#define FIX_POSE(P) s.Set##P##PoseDataFromStackedHost(F.const_pose.data(), 0, n)
#define VAR_POINT(P) s.Set##P##PointIndicesFromHost(F.point_idx.data(), n)
#define FIX_POINT(P) s.Set##P##PointDataFromStackedHost(F.const_point.data(), 0, n)
#define VAR_CALIB(P) s.Set##P##CalibIndicesFromHost(F.calib_idx.data(), n)
#define VAR_FAE_SR(P) s.Set##P##FocalAndExtraIndicesFromHost(F.calib_idx.data(), n)
#define VAR_FAE_PH(P) s.Set##P##FocalIndicesFromHost(F.calib_idx.data(), n)
#define FIX_PP(P) s.Set##P##PrincipalPointDataFromStackedHost(F.const_pp.data(), 0, n)

void set_simple_radial(caspar::GraphSolver& s, int key, const Factors& F) {
  const size_t n = F.n;
  switch (key) {
    case 7: s.SetSimpleRadialNum(n); VAR_POSE(SimpleRadial); SET_RIG(SimpleRadial); VAR_CALIB(SimpleRadial); VAR_POINT(SimpleRadial); SET_PIX(SimpleRadial); break;
    case 3: s.SetSimpleRadialFixedPoseNum(n); SET_RIG(SimpleRadialFixedPose); VAR_CALIB(SimpleRadialFixedPose); VAR_POINT(SimpleRadialFixedPose); FIX_POSE(SimpleRadialFixedPose); SET_PIX(SimpleRadialFixedPose); break;
    case 6: s.SetSimpleRadialFixedPointNum(n); VAR_POSE(SimpleRadialFixedPoint); SET_RIG(SimpleRadialFixedPoint); VAR_CALIB(SimpleRadialFixedPoint); FIX_POINT(SimpleRadialFixedPoint); SET_PIX(SimpleRadialFixedPoint); break;
    case 2: s.SetSimpleRadialFixedPoseFixedPointNum(n); SET_RIG(SimpleRadialFixedPoseFixedPoint); VAR_CALIB(SimpleRadialFixedPoseFixedPoint); FIX_POSE(SimpleRadialFixedPoseFixedPoint); FIX_POINT(SimpleRadialFixedPoseFixedPoint); SET_PIX(SimpleRadialFixedPoseFixedPoint); break;
    case 5: s.SetSimpleRadialSplitFixedPrincipalPointNum(n); VAR_POSE(SimpleRadialSplitFixedPrincipalPoint); SET_RIG(SimpleRadialSplitFixedPrincipalPoint); VAR_FAE_SR(SimpleRadialSplitFixedPrincipalPoint); VAR_POINT(SimpleRadialSplitFixedPrincipalPoint); FIX_PP(SimpleRadialSplitFixedPrincipalPoint); SET_PIX(SimpleRadialSplitFixedPrincipalPoint); break;
    case 1: s.SetSimpleRadialSplitFixedPoseFixedPrincipalPointNum(n); SET_RIG(SimpleRadialSplitFixedPoseFixedPrincipalPoint); VAR_FAE_SR(SimpleRadialSplitFixedPoseFixedPrincipalPoint); VAR_POINT(SimpleRadialSplitFixedPoseFixedPrincipalPoint); FIX_POSE(SimpleRadialSplitFixedPoseFixedPrincipalPoint); FIX_PP(SimpleRadialSplitFixedPoseFixedPrincipalPoint); SET_PIX(SimpleRadialSplitFixedPoseFixedPrincipalPoint); break;
    case 4: s.SetSimpleRadialSplitFixedPrincipalPointFixedPointNum(n); VAR_POSE(SimpleRadialSplitFixedPrincipalPointFixedPoint); SET_RIG(SimpleRadialSplitFixedPrincipalPointFixedPoint); VAR_FAE_SR(SimpleRadialSplitFixedPrincipalPointFixedPoint); FIX_PP(SimpleRadialSplitFixedPrincipalPointFixedPoint); FIX_POINT(SimpleRadialSplitFixedPrincipalPointFixedPoint); SET_PIX(SimpleRadialSplitFixedPrincipalPointFixedPoint); break;
    default: s.SetSimpleRadialSplitFixedPoseFixedPrincipalPointFixedPointNum(n); SET_RIG(SimpleRadialSplitFixedPoseFixedPrincipalPointFixedPoint); VAR_FAE_SR(SimpleRadialSplitFixedPoseFixedPrincipalPointFixedPoint); FIX_POSE(SimpleRadialSplitFixedPoseFixedPrincipalPointFixedPoint); FIX_PP(SimpleRadialSplitFixedPoseFixedPrincipalPointFixedPoint); FIX_POINT(SimpleRadialSplitFixedPoseFixedPrincipalPointFixedPoint); SET_PIX(SimpleRadialSplitFixedPoseFixedPrincipalPointFixedPoint); break;
  }
}
void set_pinhole(caspar::GraphSolver& s, int key, const Factors& F) {
  const size_t n = F.n;
  switch (key) {
    case 7: s.SetPinholeNum(n); VAR_POSE(Pinhole); SET_RIG(Pinhole); VAR_CALIB(Pinhole); VAR_POINT(Pinhole); SET_PIX(Pinhole); break;
    case 3: s.SetPinholeFixedPoseNum(n); SET_RIG(PinholeFixedPose); VAR_CALIB(PinholeFixedPose); VAR_POINT(PinholeFixedPose); FIX_POSE(PinholeFixedPose); SET_PIX(PinholeFixedPose); break;
    case 6: s.SetPinholeFixedPointNum(n); VAR_POSE(PinholeFixedPoint); SET_RIG(PinholeFixedPoint); VAR_CALIB(PinholeFixedPoint); FIX_POINT(PinholeFixedPoint); SET_PIX(PinholeFixedPoint); break;
    case 2: s.SetPinholeFixedPoseFixedPointNum(n); SET_RIG(PinholeFixedPoseFixedPoint); VAR_CALIB(PinholeFixedPoseFixedPoint); FIX_POSE(PinholeFixedPoseFixedPoint); FIX_POINT(PinholeFixedPoseFixedPoint); SET_PIX(PinholeFixedPoseFixedPoint); break;
    case 5: s.SetPinholeSplitFixedPrincipalPointNum(n); VAR_POSE(PinholeSplitFixedPrincipalPoint); SET_RIG(PinholeSplitFixedPrincipalPoint); VAR_FAE_PH(PinholeSplitFixedPrincipalPoint); VAR_POINT(PinholeSplitFixedPrincipalPoint); FIX_PP(PinholeSplitFixedPrincipalPoint); SET_PIX(PinholeSplitFixedPrincipalPoint); break;
    case 1: s.SetPinholeSplitFixedPoseFixedPrincipalPointNum(n); SET_RIG(PinholeSplitFixedPoseFixedPrincipalPoint); VAR_FAE_PH(PinholeSplitFixedPoseFixedPrincipalPoint); VAR_POINT(PinholeSplitFixedPoseFixedPrincipalPoint); FIX_POSE(PinholeSplitFixedPoseFixedPrincipalPoint); FIX_PP(PinholeSplitFixedPoseFixedPrincipalPoint); SET_PIX(PinholeSplitFixedPoseFixedPrincipalPoint); break;
    case 4: s.SetPinholeSplitFixedPrincipalPointFixedPointNum(n); VAR_POSE(PinholeSplitFixedPrincipalPointFixedPoint); SET_RIG(PinholeSplitFixedPrincipalPointFixedPoint); VAR_FAE_PH(PinholeSplitFixedPrincipalPointFixedPoint); FIX_PP(PinholeSplitFixedPrincipalPointFixedPoint); FIX_POINT(PinholeSplitFixedPrincipalPointFixedPoint); SET_PIX(PinholeSplitFixedPrincipalPointFixedPoint); break;
    default: s.SetPinholeSplitFixedPoseFixedPrincipalPointFixedPointNum(n); SET_RIG(PinholeSplitFixedPoseFixedPrincipalPointFixedPoint); VAR_FAE_PH(PinholeSplitFixedPoseFixedPrincipalPointFixedPoint); FIX_POSE(PinholeSplitFixedPoseFixedPrincipalPointFixedPoint); FIX_PP(PinholeSplitFixedPoseFixedPrincipalPointFixedPoint); FIX_POINT(PinholeSplitFixedPoseFixedPrincipalPointFixedPoint); SET_PIX(PinholeSplitFixedPoseFixedPrincipalPointFixedPoint); break;
  }
}

}  // namespace

struct caspar_ref_result {
  int iterations, exit_reason;
  double initial_cost, final_cost;   // 1/2 sum |r|^2 as Caspar scores it
  double solve_ms;                   // wall time of GraphSolver::solve (device-synchronous)
  double setup_ms;                   // flattening + H2D (constructor, setters, finish_indices)
  int num_residuals;
};

extern "C" int caspar_ref_solve(const b200ba_options* o, b200ba_problem* p, int solver_iter_max, caspar_ref_result* out,
                                char* err, int errlen) {
  auto fail = [&](const std::string& m) { if (err && errlen > 0) { strncpy(err, m.c_str(), errlen - 1); err[errlen - 1] = 0; } return -1; };
  try {
    const auto t0 = std::chrono::steady_clock::now();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("no CUDA device");
    if (!o->refine_focal_length || !o->refine_extra_params) return fail("harness covers refine_focal_length == refine_extra_params == true");
    const int NI = p->num_poses, NC = p->num_cameras;
    const int64_t NP = p->num_points, NO = p->num_observations;
    ModelPool pool[2];   // 0 = SIMPLE_RADIAL (model id 2), 1 = PINHOLE (model id 1)
    for (auto& m : pool) { m.pose_of_image.assign(NI, -1); m.calib_of_camera.assign(NC, -1); }
    auto model_slot = [&](int cam) { const int id = p->camera_model_id[cam]; return id == 2 ? 0 : (id == 1 ? 1 : -1); };
    auto pose_var = [&](int i) { return o->refine_rig_from_world && !(p->pose_constant && p->pose_constant[i]); };
    auto cam_var = [&](int c) { return !(p->camera_constant && p->camera_constant[c]); };
    auto point_var = [&](int64_t k) { return o->refine_points3D && !(p->point_constant && p->point_constant[k]); };
    std::vector<float> points(3 * (size_t)NP);
    for (size_t i = 0; i < points.size(); ++i) points[i] = (float)p->points[i];
    const float identity[7] = {0, 0, 0, 1, 0, 0, 0};
    size_t nres = 0;
    for (int64_t k = 0; k < NO; ++k) {
      const int img = p->obs_pose_idx[k], cam = p->obs_camera_idx[k];
      const int64_t pt = p->obs_point_idx[k];
      const int ms = model_slot(cam);
      if (ms < 0) return fail("Caspar has SIMPLE_RADIAL and PINHOLE only");
      ModelPool& M = pool[ms];
      const bool pv = pose_var(img), cv = cam_var(cam), ptv = point_var(pt);
      if (!cv) return fail("harness covers variable focal_and_extra only");
      const bool ppv = cv && o->refine_principal_point;
      if (M.calib_of_camera[cam] < 0) {
        M.calib_of_camera[cam] = (int)M.node_camera.size(); M.node_camera.push_back(cam);
        const double* q = p->camera_params + p->camera_param_offset[cam];
        if (ms == 0) { M.fae.push_back((float)q[0]); M.fae.push_back((float)q[3]); M.pp.push_back((float)q[1]); M.pp.push_back((float)q[2]); }
        else { M.fae.push_back((float)q[0]); M.fae.push_back((float)q[1]); M.pp.push_back((float)q[2]); M.pp.push_back((float)q[3]); }
      }
      const int calib = M.calib_of_camera[cam];
      Factors& F = M.f[(pv ? 4 : 0) | (ppv ? 2 : 0) | (ptv ? 1 : 0)];
      F.rig.insert(F.rig.end(), identity, identity + 7);
      if (pv) {
        if (M.pose_of_image[img] < 0) {
          M.pose_of_image[img] = (int)M.node_image.size(); M.node_image.push_back(img);
          for (int c = 0; c < 7; ++c) M.pose_data.push_back((float)p->poses[7 * (size_t)img + c]);
        }
        F.pose_idx.push_back((unsigned)M.pose_of_image[img]);
      } else {
        for (int c = 0; c < 7; ++c) F.const_pose.push_back((float)p->poses[7 * (size_t)img + c]);
      }
      F.calib_idx.push_back((unsigned)calib);
      if (!ppv) { F.const_pp.push_back(M.pp[2 * calib]); F.const_pp.push_back(M.pp[2 * calib + 1]); }
      if (ptv) F.point_idx.push_back((unsigned)pt);
      else for (int c = 0; c < 3; ++c) F.const_point.push_back(points[3 * (size_t)pt + c]);
      F.pixels.push_back((float)p->obs_xy[2 * k]); F.pixels.push_back((float)p->obs_xy[2 * k + 1]);
      F.n += 1; nres += 2;
    }
    caspar::SolverParams<float> params;   // CasparBundleAdjustmentOptions defaults
    params.solver_iter_max = solver_iter_max > 0 ? solver_iter_max : 200;
    params.pcg_iter_max = 20; params.diag_init = 1.0f; params.diag_min = 1e-12f; params.diag_scaling_up = 2.0f;
    params.diag_scaling_down = 0.333333f; params.diag_exit_value = 1e3f; params.score_exit_value = 0.0f;
    params.pcg_rel_error_exit = 1e-4f; params.pcg_rel_score_exit = -1.0f; params.pcg_rel_decrease_min = -1.0f;
    params.solver_rel_decrease_min = 1.0f;
    int dev = 0;
    if (o->gpu_index >= 0) dev = o->gpu_index; else cudaGetDevice(&dev);
    const ModelPool &SR = pool[0], &PH = pool[1];
    // constructor argument order: caspar_model_adapter.h:922-984 (node pools alphabetical, then the factor counts)
    caspar::GraphSolver solver(
        params, PH.node_camera.size(), PH.node_camera.size(), PH.node_image.size(), PH.node_camera.size(), (size_t)NP,
        SR.node_camera.size(), SR.node_camera.size(), SR.node_image.size(), SR.node_camera.size(),
        SR.f[7].n, SR.f[3].n, SR.f[6].n, SR.f[2].n, PH.f[7].n, PH.f[3].n, PH.f[6].n, PH.f[2].n,
        /* simple_radial_split */ 0, SR.f[5].n, 0, SR.f[1].n, 0, 0, SR.f[4].n, 0, 0, SR.f[0].n, 0,
        /* pinhole_split */ 0, PH.f[5].n, 0, PH.f[1].n, 0, 0, PH.f[4].n, 0, 0, PH.f[0].n, 0, (size_t)dev);
    if (NP > 0) solver.SetPointNodesFromStackedHost(points.data(), 0, (size_t)NP);
    auto merged = [](const ModelPool& M) { return M.f[7].n + M.f[3].n + M.f[6].n + M.f[2].n > 0; };
    auto calib4 = [](const ModelPool& M) { std::vector<float> c(4 * M.node_camera.size()); for (size_t i = 0; i < M.node_camera.size(); ++i) { c[4 * i] = M.fae[2 * i]; c[4 * i + 1] = M.fae[2 * i + 1]; c[4 * i + 2] = M.pp[2 * i]; c[4 * i + 3] = M.pp[2 * i + 1]; } return c; };
    if (!SR.node_image.empty()) solver.SetSimpleRadialPoseNodesFromStackedHost(SR.pose_data.data(), 0, SR.node_image.size());
    if (!PH.node_image.empty()) solver.SetPinholePoseNodesFromStackedHost(PH.pose_data.data(), 0, PH.node_image.size());
    if (!SR.node_camera.empty()) {
      solver.SetSimpleRadialFocalAndExtraNodesFromStackedHost(SR.fae.data(), 0, SR.node_camera.size());
      solver.SetSimpleRadialPrincipalPointNodesFromStackedHost(SR.pp.data(), 0, SR.node_camera.size());
      if (merged(SR)) { auto c = calib4(SR); solver.SetSimpleRadialCalibNodesFromStackedHost(c.data(), 0, SR.node_camera.size()); }
    }
    if (!PH.node_camera.empty()) {
      solver.SetPinholeFocalNodesFromStackedHost(PH.fae.data(), 0, PH.node_camera.size());
      solver.SetPinholePrincipalPointNodesFromStackedHost(PH.pp.data(), 0, PH.node_camera.size());
      if (merged(PH)) { auto c = calib4(PH); solver.SetPinholeCalibNodesFromStackedHost(c.data(), 0, PH.node_camera.size()); }
    }
    for (int key = 0; key < 8; ++key) {
      if (SR.f[key].n) set_simple_radial(solver, key, SR.f[key]);
      if (PH.f[key].n) set_pinhole(solver, key, PH.f[key]);
    }
    solver.finish_indices();
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    caspar::SolveResult res = solver.solve(false, false);
    cudaDeviceSynchronize();
    const auto t2 = std::chrono::steady_clock::now();
    // read back (ReadSolverResults / WriteResultsToReconstruction, :674-801): variable blocks only
    if (NP > 0) {
      solver.GetPointNodesToStackedHost(points.data(), 0, (size_t)NP);
      for (int64_t k = 0; k < NP; ++k) if (point_var(k)) for (int c = 0; c < 3; ++c) p->points[3 * k + c] = points[3 * (size_t)k + c];
    }
    for (int ms = 0; ms < 2; ++ms) {
      ModelPool& M = pool[ms];
      if (!M.node_image.empty()) {
        if (ms == 0) solver.GetSimpleRadialPoseNodesToStackedHost(M.pose_data.data(), 0, M.node_image.size());
        else solver.GetPinholePoseNodesToStackedHost(M.pose_data.data(), 0, M.node_image.size());
        for (size_t n = 0; n < M.node_image.size(); ++n) for (int c = 0; c < 7; ++c) p->poses[7 * (size_t)M.node_image[n] + c] = M.pose_data[7 * n + c];
      }
      if (!M.node_camera.empty()) {
        std::vector<float> fae(M.fae.size()), pp(M.pp.size());
        if (merged(M)) {
          std::vector<float> c(4 * M.node_camera.size());
          if (ms == 0) solver.GetSimpleRadialCalibNodesToStackedHost(c.data(), 0, M.node_camera.size());
          else solver.GetPinholeCalibNodesToStackedHost(c.data(), 0, M.node_camera.size());
          for (size_t i = 0; i < M.node_camera.size(); ++i) { fae[2 * i] = c[4 * i]; fae[2 * i + 1] = c[4 * i + 1]; pp[2 * i] = c[4 * i + 2]; pp[2 * i + 1] = c[4 * i + 3]; }
        } else {
          if (ms == 0) solver.GetSimpleRadialFocalAndExtraNodesToStackedHost(fae.data(), 0, M.node_camera.size());
          else solver.GetPinholeFocalNodesToStackedHost(fae.data(), 0, M.node_camera.size());
          pp = M.pp;
        }
        for (size_t i = 0; i < M.node_camera.size(); ++i) {
          double* q = p->camera_params + p->camera_param_offset[M.node_camera[i]];
          if (ms == 0) { q[0] = fae[2 * i]; q[3] = fae[2 * i + 1]; if (o->refine_principal_point) { q[1] = pp[2 * i]; q[2] = pp[2 * i + 1]; } }
          else { q[0] = fae[2 * i]; q[1] = fae[2 * i + 1]; if (o->refine_principal_point) { q[2] = pp[2 * i]; q[3] = pp[2 * i + 1]; } }
        }
      }
    }
    out->iterations = res.iteration_count; out->exit_reason = (int)res.exit_reason;
    out->initial_cost = 0.0 /* SolveResult::initial_score is never written by the generated solver */; out->final_cost = res.final_score;
    out->solve_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    out->setup_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    out->num_residuals = (int)nres;
    return 0;
  } catch (const std::exception& e) {
    return fail(e.what());
  }
}
