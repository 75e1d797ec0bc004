// mvs_fusion.cu — stereo fusion behind include/b200_mvs_fusion.h (host code only; a .cu so that the one build recipe
// picks it up).  Restates StereoFusion::Run / Fuse (src/colmap/mvs/fusion.cc:131-545) in its single-threaded schedule.
// fp32 arithmetic where the reference uses fp32 (projection, errors, back-projection); two documented deviations at
// rounding level: the inverse projection is composed in closed form ([R^T K^-1 | -R^T T], in double, rounded once)
// instead of a 4x4 fp32 inverse, and visibility lists are sorted.
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200_mvs_fusion.h"

static thread_local std::string g_fuse_error;
static int fuse_fail(int code, const std::string& msg) { g_fuse_error = msg; return code; }

struct b200fuse_result {
  std::vector<float> xyz, normal;
  std::vector<uint8_t> rgb;
  std::vector<int64_t> vis_offset{0};
  std::vector<int32_t> vis;
};

namespace {

// Percentile(elems, 50) (math/math.h:205-234)
template <typename T>
double median(std::vector<T>& v) {
  const double idx = 0.5 * (double)(v.size() - 1);
  const double lo_d = floor(idx), hi_d = ceil(idx);
  const size_t lo = (size_t)lo_d, hi = (size_t)hi_d;
  std::nth_element(v.begin(), v.begin() + hi, v.end());
  const double right = (double)v[hi];
  if (lo == hi) return right;
  const double left = (double)*std::max_element(v.begin(), v.begin() + hi);
  return (hi_d - idx) * left + (idx - lo_d) * right;
}

struct View {
  float P[12], invP[12], invR[9];
  float sx, sy;   // depth-map size / image size
};

// ComposeProjectionMatrix (mvs/image.cc:114-123): P = K [R | T], fp32
void compose_P(const float* K, const float* R, const float* T, float* P) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      const float m0 = c < 3 ? R[c] : T[0], m1 = c < 3 ? R[3 + c] : T[1], m2 = c < 3 ? R[6 + c] : T[2];
      P[4 * r + c] = (K[3 * r] * m0 + K[3 * r + 1] * m1) + K[3 * r + 2] * m2;
    }
}
// inverse of [K R | K T; 0 0 0 1], top three rows: [R^T K^-1 | -R^T T]
void compose_invP(const float* Kf, const float* R, const float* T, float* invP) {
  double K[9];
  for (int i = 0; i < 9; ++i) K[i] = Kf[i];
  const double det = K[0] * (K[4] * K[8] - K[5] * K[7]) - K[1] * (K[3] * K[8] - K[5] * K[6]) + K[2] * (K[3] * K[7] - K[4] * K[6]);
  double iK[9] = {(K[4] * K[8] - K[5] * K[7]) / det, (K[2] * K[7] - K[1] * K[8]) / det, (K[1] * K[5] - K[2] * K[4]) / det,
                  (K[5] * K[6] - K[3] * K[8]) / det, (K[0] * K[8] - K[2] * K[6]) / det, (K[2] * K[3] - K[0] * K[5]) / det,
                  (K[3] * K[7] - K[4] * K[6]) / det, (K[1] * K[6] - K[0] * K[7]) / det, (K[0] * K[4] - K[1] * K[3]) / det};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      invP[4 * r + c] = (float)(((double)R[r] * iK[c] + (double)R[3 + r] * iK[3 + c]) + (double)R[6 + r] * iK[6 + c]);
    invP[4 * r + 3] = (float)(-(((double)R[r] * T[0] + (double)R[3 + r] * T[1]) + (double)R[6 + r] * T[2]));
  }
}

}  // namespace

extern "C" {

const char* b200fuse_last_error(void) { return g_fuse_error.c_str(); }

void b200fuse_options_init(b200fuse_options* o) {
  o->min_num_pixels = 5; o->max_num_pixels = 10000; o->max_traversal_depth = 100;
  o->max_reproj_error = 2.0f; o->max_depth_error = 0.01f; o->max_normal_error = 10.0f;
  for (int k = 0; k < 3; ++k) { o->bbox_min[k] = -FLT_MAX; o->bbox_max[k] = FLT_MAX; }
}

int b200fuse_options_check(const b200fuse_options* o) {
  return o && o->min_num_pixels >= 0 && o->min_num_pixels <= o->max_num_pixels && o->max_traversal_depth > 0 &&
         o->max_reproj_error >= 0 && o->max_depth_error >= 0 && o->max_normal_error >= 0;
}

int b200fuse_run(const b200fuse_options* o, int n, const b200fuse_image* images, const int32_t* overlap, int max_overlap,
                 b200fuse_result_t* out) {
  if (!o || !images || !out || n < 0 || max_overlap < 0) return fuse_fail(-1, "null / negative argument");
  if (!b200fuse_options_check(o)) return fuse_fail(-2, "StereoFusionOptions::Check failed");
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < max_overlap; ++k) { const int v = overlap[(size_t)i * max_overlap + k]; if (v < -1 || v >= n) return fuse_fail(-3, "overlap list references an image outside the model"); }
  const float max_sq_reproj = (float)(o->max_reproj_error * o->max_reproj_error);
  const float min_cos_normal = (float)cos(o->max_normal_error * 0.0174532925199432954743716805978692718781530857086181640625);
  const float max_depth_error = (float)o->max_depth_error;
  std::vector<View> V(n);
  std::vector<std::vector<uint8_t>> mask(n);
  std::vector<char> used(n, 0), fused(n, 0);
  for (int i = 0; i < n; ++i) {
    const b200fuse_image& im = images[i];
    if (!im.used) continue;
    if (!im.K || !im.R || !im.T || !im.depth || !im.normal || !im.rgb || im.map_width <= 0 || im.map_height <= 0 ||
        im.image_width <= 0 || im.image_height <= 0 || im.bitmap_width <= 0 || im.bitmap_height <= 0)
      return fuse_fail(-4, "used image with missing inputs");
    used[i] = 1;
    View& v = V[i];
    v.sx = (float)im.map_width / (float)im.image_width; v.sy = (float)im.map_height / (float)im.image_height;
    float K[9];
    memcpy(K, im.K, sizeof(K));
    K[0] *= v.sx; K[2] *= v.sx; K[4] *= v.sy; K[5] *= v.sy;
    compose_P(K, im.R, im.T, v.P);
    compose_invP(K, im.R, im.T, v.invP);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) v.invR[3 * r + c] = im.R[3 * c + r];
    mask[i].assign((size_t)im.map_width * im.map_height, 0);
    if (im.mask) for (size_t k = 0; k < mask[i].size(); ++k) mask[i][k] = im.mask[k] ? 1 : 0;
  }
  b200fuse_result* res = new b200fuse_result();
  struct Item { int image, row, col, depth; };
  std::vector<Item> queue;
  std::vector<float> px, py, pz, nx, ny, nz;
  std::vector<uint8_t> cr, cg, cb;
  std::vector<int32_t> vis;

  auto fuse = [&](int image0, int row0, int col0) {
    queue.clear(); px.clear(); py.clear(); pz.clear(); nx.clear(); ny.clear(); nz.clear(); cr.clear(); cg.clear(); cb.clear(); vis.clear();
    queue.push_back({image0, row0, col0, 0});
    float ref_point[4] = {0, 0, 0, 0}, ref_normal[3] = {0, 0, 0};
    while (!queue.empty()) {
      const Item d = queue.back();
      queue.pop_back();
      const b200fuse_image& im = images[d.image];
      const View& v = V[d.image];
      const size_t pix = (size_t)d.row * im.map_width + d.col;
      if (mask[d.image][pix]) continue;
      const float depth = im.depth[pix];
      if (depth <= 0.0f) continue;
      if (d.depth > 0) {
        float proj[3];
        for (int r = 0; r < 3; ++r)
          proj[r] = ((v.P[4 * r] * ref_point[0] + v.P[4 * r + 1] * ref_point[1]) + v.P[4 * r + 2] * ref_point[2]) + v.P[4 * r + 3] * ref_point[3];
        const float depth_error = fabsf((proj[2] - depth) / depth);
        if (depth_error > max_depth_error) continue;
        const float col_diff = proj[0] / proj[2] - (float)d.col, row_diff = proj[1] / proj[2] - (float)d.row;
        if (col_diff * col_diff + row_diff * row_diff > max_sq_reproj) continue;
      }
      const size_t plane = (size_t)im.map_width * im.map_height;
      const float n0 = im.normal[pix], n1 = im.normal[plane + pix], n2 = im.normal[2 * plane + pix];
      float normal[3];
      for (int r = 0; r < 3; ++r) normal[r] = (v.invR[3 * r] * n0 + v.invR[3 * r + 1] * n1) + v.invR[3 * r + 2] * n2;
      if (d.depth > 0) {
        const float c = (ref_normal[0] * normal[0] + ref_normal[1] * normal[1]) + ref_normal[2] * normal[2];
        if (c < min_cos_normal) continue;
      }
      const float h[4] = {(float)d.col * depth, (float)d.row * depth, depth, 1.0f};
      float xyz[3];
      for (int r = 0; r < 3; ++r) xyz[r] = ((v.invP[4 * r] * h[0] + v.invP[4 * r + 1] * h[1]) + v.invP[4 * r + 2] * h[2]) + v.invP[4 * r + 3] * h[3];
      // colour: nearest neighbour in the bitmap (Bitmap::InterpolateNearestNeighbor), black outside
      uint8_t color[3] = {0, 0, 0};
      {
        const int xx = (int)round((double)((float)d.col / v.sx)), yy = (int)round((double)((float)d.row / v.sy));
        if (xx >= 0 && yy >= 0 && xx < im.bitmap_width && yy < im.bitmap_height)
          memcpy(color, im.rgb + 3 * ((size_t)yy * im.bitmap_width + xx), 3);
      }
      mask[d.image][pix] = 1;
      if (xyz[0] < o->bbox_min[0] || xyz[1] < o->bbox_min[1] || xyz[2] < o->bbox_min[2] || xyz[0] > o->bbox_max[0] ||
          xyz[1] > o->bbox_max[1] || xyz[2] > o->bbox_max[2])
        continue;
      px.push_back(xyz[0]); py.push_back(xyz[1]); pz.push_back(xyz[2]);
      nx.push_back(normal[0]); ny.push_back(normal[1]); nz.push_back(normal[2]);
      cr.push_back(color[0]); cg.push_back(color[1]); cb.push_back(color[2]);
      if (std::find(vis.begin(), vis.end(), d.image) == vis.end()) vis.push_back(d.image);
      if (d.depth == 0) {
        ref_point[0] = xyz[0]; ref_point[1] = xyz[1]; ref_point[2] = xyz[2]; ref_point[3] = 1.0f;
        memcpy(ref_normal, normal, sizeof(ref_normal));
      }
      if (px.size() >= (size_t)o->max_num_pixels) break;
      if (d.depth >= o->max_traversal_depth - 1) continue;
      for (int k = 0; k < max_overlap; ++k) {
        const int next = overlap[(size_t)d.image * max_overlap + k];
        if (next < 0) break;
        if (!used[next] || fused[next]) continue;
        const View& w = V[next];
        float np[3];
        for (int r = 0; r < 3; ++r) np[r] = ((w.P[4 * r] * xyz[0] + w.P[4 * r + 1] * xyz[1]) + w.P[4 * r + 2] * xyz[2]) + w.P[4 * r + 3];
        const int next_col = (int)roundf(np[0] / np[2]), next_row = (int)roundf(np[1] / np[2]);
        if (next_col < 0 || next_row < 0 || next_col >= images[next].map_width || next_row >= images[next].map_height) continue;
        queue.push_back({next, next_row, next_col, d.depth + 1});
      }
    }
    if (px.size() < (size_t)o->min_num_pixels || px.empty()) return;
    const float fnx = (float)median(nx), fny = (float)median(ny), fnz = (float)median(nz);
    const float norm = sqrtf((fnx * fnx + fny * fny) + fnz * fnz);
    if (norm < FLT_EPSILON) return;
    res->xyz.push_back((float)median(px)); res->xyz.push_back((float)median(py)); res->xyz.push_back((float)median(pz));
    res->normal.push_back(fnx / norm); res->normal.push_back(fny / norm); res->normal.push_back(fnz / norm);
    auto to_u8 = [](double m) { const float f = roundf((float)m); return (uint8_t)std::min(255.0f, std::max(0.0f, f)); };
    res->rgb.push_back(to_u8(median(cr))); res->rgb.push_back(to_u8(median(cg))); res->rgb.push_back(to_u8(median(cb)));
    std::sort(vis.begin(), vis.end());
    res->vis.insert(res->vis.end(), vis.begin(), vis.end());
    res->vis_offset.push_back((int64_t)res->vis.size());
  };

  // image order: start at 0, then FindNextImage (fusion.cc:48-72)
  for (int image = n > 0 ? 0 : -1; image >= 0;) {
    if (used[image]) {
      const b200fuse_image& im = images[image];
      for (int row = 0; row < im.map_height; ++row)
        for (int col = 0; col < im.map_width; ++col)
          if (!mask[image][(size_t)row * im.map_width + col]) fuse(image, row, col);
    }
    fused[image] = 1;
    int next = -1;
    for (int k = 0; k < max_overlap && next < 0; ++k) {
      const int c = overlap[(size_t)image * max_overlap + k];
      if (c < 0) break;
      if (used[c] && !fused[c]) next = c;
    }
    for (int i = 0; i < n && next < 0; ++i) if (used[i] && !fused[i]) next = i;
    image = next;
  }
  *out = res;
  return 0;
}

int64_t b200fuse_num_points(b200fuse_result_t r) { return r ? (int64_t)(r->xyz.size() / 3) : -1; }
int64_t b200fuse_num_visibility(b200fuse_result_t r) { return r ? (int64_t)r->vis.size() : -1; }
int b200fuse_get(b200fuse_result_t r, float* xyz, float* normal, uint8_t* rgb, int64_t* vis_offset, int32_t* vis) {
  if (!r) return fuse_fail(-1, "null result");
  if (xyz) memcpy(xyz, r->xyz.data(), sizeof(float) * r->xyz.size());
  if (normal) memcpy(normal, r->normal.data(), sizeof(float) * r->normal.size());
  if (rgb) memcpy(rgb, r->rgb.data(), r->rgb.size());
  if (vis_offset) memcpy(vis_offset, r->vis_offset.data(), sizeof(int64_t) * r->vis_offset.size());
  if (vis) memcpy(vis, r->vis.data(), sizeof(int32_t) * r->vis.size());
  return 0;
}
void b200fuse_free(b200fuse_result_t r) { delete r; }

int b200fuse_write_visibility(const char* path, int64_t num_points, const int64_t* vis_offset, const int32_t* vis) {
  FILE* f = fopen(path, "wb");
  if (!f) return fuse_fail(-10, std::string("cannot open ") + path);
  const uint64_t n = (uint64_t)num_points;
  fwrite(&n, 8, 1, f);
  for (int64_t i = 0; i < num_points; ++i) {
    const uint32_t m = (uint32_t)(vis_offset[i + 1] - vis_offset[i]);
    fwrite(&m, 4, 1, f);
    for (int64_t k = vis_offset[i]; k < vis_offset[i + 1]; ++k) { const uint32_t v = (uint32_t)vis[k]; fwrite(&v, 4, 1, f); }
  }
  fclose(f);
  return 0;
}

int b200fuse_read_visibility(const char* path, int64_t num_points, int64_t* vis_offset, int32_t* vis, int64_t* num_visibility) {
  FILE* f = fopen(path, "rb");
  if (!f) return fuse_fail(-10, std::string("cannot open ") + path);
  uint64_t n = 0;
  if (fread(&n, 8, 1, f) != 1) { fclose(f); return fuse_fail(-13, "truncated visibility file"); }
  if ((int64_t)n != num_points) { fclose(f); return fuse_fail(-15, "visibility file holds a different number of points"); }
  int64_t total = 0;
  if (vis_offset) vis_offset[0] = 0;
  for (int64_t i = 0; i < num_points; ++i) {
    uint32_t m;
    if (fread(&m, 4, 1, f) != 1) { fclose(f); return fuse_fail(-13, "truncated visibility file"); }
    for (uint32_t k = 0; k < m; ++k) {
      uint32_t v;
      if (fread(&v, 4, 1, f) != 1) { fclose(f); return fuse_fail(-13, "truncated visibility file"); }
      if (vis) vis[total + k] = (int32_t)v;
    }
    total += m;
    if (vis_offset) vis_offset[i + 1] = total;
  }
  fclose(f);
  if (num_visibility) *num_visibility = total;
  return 0;
}

}  // extern "C"
