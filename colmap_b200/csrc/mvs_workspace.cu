// mvs_workspace.cu — host-side neighbours of the PatchMatch sweep behind include/b200_mvs_workspace.h (no device code;
// the file is a .cu only so that the one build recipe picks it up).
//
// Reference behaviour restated here: the map file formats (mvs/mat.cc:41-66, mvs/consistency_graph.cc:66-139), the
// sparse-model statistics of mvs/model.cc:120-283 (depth ranges, shared points, triangulation angles, overlapping
// images) and the patch-match.cfg reader of PatchMatchController::ReadProblems (mvs/patch_match.cc:240-372).
// Little-endian hosts only (the reference byte-swaps on big-endian machines; a B200 host is x86-64 / aarch64 LE).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/b200_mvs_workspace.h"

static thread_local std::string g_ws_error;
static int ws_fail(int code, const std::string& msg) { g_ws_error = msg; return code; }

namespace {

struct File {
  FILE* f = nullptr;
  File(const char* path, const char* mode) { f = fopen(path, mode); }
  ~File() { if (f) fclose(f); }
};

// "W&H&D&" — formatted extraction like `file >> w >> c >> h >> c >> d >> c` (leading whitespace skipped)
bool read_header(FILE* f, long long v[3]) {
  for (int k = 0; k < 3; ++k) {
    if (fscanf(f, "%lld", &v[k]) != 1) return false;
    int c;
    do { c = fgetc(f); } while (c == ' ' || c == '\n' || c == '\t' || c == '\r');
    if (c == EOF) return false;
  }
  return true;
}

// Percentile (math/math.h:205-224): linear interpolation between the order statistics around p/100 (n-1)
double percentile(std::vector<float>& v, double p) {
  const double idx = p / 100.0 * (double)(v.size() - 1);
  const double lo_d = floor(idx), hi_d = ceil(idx);
  const size_t lo = (size_t)lo_d, hi = (size_t)hi_d;
  std::nth_element(v.begin(), v.begin() + hi, v.end());
  const double right = v[hi];
  if (lo == hi) return right;
  const double left = *std::max_element(v.begin(), v.begin() + hi);
  return (hi_d - idx) * left + (idx - lo_d) * right;
}

// ComputeProjectionCenter (mvs/image.cc:97-104) in fp32: C = -R^T T
void projection_center(const float* R, const float* T, double C[3]) {
  for (int c = 0; c < 3; ++c) {
    const float v = -(R[c] * T[0] + R[3 + c] * T[1] + R[6 + c] * T[2]);
    C[c] = (double)v;
  }
}

// CalculateTriangulationAngle (geometry/triangulation.cc:217-250)
double triangulation_angle(const double* c1, const double* c2, const double* X) {
  double a[3], b[3];
  for (int k = 0; k < 3; ++k) { a[k] = X[k] - c1[k]; b[k] = X[k] - c2[k]; }
  const double n1 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], n2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  double angle = 0.0;
  if (n1 != 0.0 && n2 != 0.0) {
    double c = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / sqrt(n1 * n2);
    c = std::min(1.0, std::max(-1.0, c));
    angle = acos(c);
  }
  return std::min(angle, M_PI - angle);
}

int check_model(const b200ws_model* m) {
  if (!m || m->num_images < 0 || m->num_points < 0) return ws_fail(-1, "invalid model");
  for (int64_t p = 0; p < m->num_points; ++p)
    for (int64_t j = m->track_offset[p]; j < m->track_offset[p + 1]; ++j)
      if (m->track[j] < 0 || m->track[j] >= m->num_images) return ws_fail(-2, "track references an image outside the model");
  return 0;
}

// pairwise statistics of the tracks: count of shared points and (optionally) all triangulation angles per image pair
void pair_statistics(const b200ws_model* m, std::vector<int32_t>* counts, std::vector<std::map<int, std::vector<float>>>* angles) {
  const int n = m->num_images;
  std::vector<double> centers;
  if (angles) {
    centers.resize(3 * (size_t)n);
    for (int i = 0; i < n; ++i) projection_center(m->R + 9 * (size_t)i, m->T + 3 * (size_t)i, centers.data() + 3 * (size_t)i);
    angles->assign(n, {});
  }
  if (counts) counts->assign((size_t)n * n, 0);
  for (int64_t p = 0; p < m->num_points; ++p) {
    const int32_t* tr = m->track + m->track_offset[p];
    const int64_t len = m->track_offset[p + 1] - m->track_offset[p];
    const double X[3] = {m->xyz[3 * p], m->xyz[3 * p + 1], m->xyz[3 * p + 2]};
    for (int64_t i = 0; i < len; ++i)
      for (int64_t j = 0; j < i; ++j) {
        const int a = tr[i], b = tr[j];
        if (a == b) continue;
        if (counts) { (*counts)[(size_t)a * n + b] += 1; (*counts)[(size_t)b * n + a] += 1; }
        if (angles) {
          const float ang = (float)triangulation_angle(centers.data() + 3 * (size_t)a, centers.data() + 3 * (size_t)b, X);
          (*angles)[a][b].push_back(ang);
          (*angles)[b][a].push_back(ang);
        }
      }
  }
}

// the ordered candidate list of one reference image (shared-point count, descending; ties by ascending index)
std::vector<int> ranked_sources(int ref, int n, const std::vector<int32_t>& counts, const std::vector<float>& tri75,
                                float min_angle_rad, size_t max_num) {
  std::vector<std::pair<int, int>> cand;   // (image, count), visited in ascending image order like std::map iteration
  for (int j = 0; j < n; ++j) {
    const int c = counts[(size_t)ref * n + j];
    if (c > 0 && tri75[(size_t)ref * n + j] >= min_angle_rad) cand.emplace_back(j, c);
  }
  std::stable_sort(cand.begin(), cand.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.second > y.second; });
  std::vector<int> out;
  for (size_t i = 0; i < cand.size() && i < max_num; ++i) out.push_back(cand[i].first);
  return out;
}

void tri_percentile_dense(const b200ws_model* m, float p, std::vector<float>* out, std::vector<int32_t>* counts) {
  const int n = m->num_images;
  std::vector<std::map<int, std::vector<float>>> all;
  pair_statistics(m, counts, &all);
  out->assign((size_t)n * n, -1.0f);
  for (int i = 0; i < n; ++i)
    for (auto& kv : all[i]) (*out)[(size_t)i * n + kv.first] = (float)percentile(kv.second, p);
}

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

}  // namespace

extern "C" {

const char* b200ws_last_error(void) { return g_ws_error.c_str(); }

int b200ws_mat_read_header(const char* path, int* width, int* height, int* depth) {
  File file(path, "rb");
  if (!file.f) return ws_fail(-10, std::string("cannot open ") + path);
  long long v[3];
  if (!read_header(file.f, v)) return ws_fail(-11, std::string("malformed map header in ") + path);
  if (v[0] <= 0 || v[1] <= 0 || v[2] <= 0) return ws_fail(-11, std::string("non-positive map size in ") + path);
  *width = (int)v[0]; *height = (int)v[1]; *depth = (int)v[2];
  return 0;
}

int b200ws_mat_read(const char* path, float* data, size_t capacity) {
  File file(path, "rb");
  if (!file.f) return ws_fail(-10, std::string("cannot open ") + path);
  long long v[3];
  if (!read_header(file.f, v) || v[0] <= 0 || v[1] <= 0 || v[2] <= 0) return ws_fail(-11, std::string("malformed map header in ") + path);
  const size_t n = (size_t)v[0] * (size_t)v[1] * (size_t)v[2];
  if (capacity < n) return ws_fail(-12, "buffer too small for the map");
  if (fread(data, sizeof(float), n, file.f) != n) return ws_fail(-13, std::string("truncated map file ") + path);
  return 0;
}

int b200ws_mat_write(const char* path, const float* data, int width, int height, int depth) {
  if (width <= 0 || height <= 0 || depth <= 0) return ws_fail(-11, "non-positive map size");
  File file(path, "wb");
  if (!file.f) return ws_fail(-10, std::string("cannot open ") + path);
  fprintf(file.f, "%d&%d&%d&", width, height, depth);
  const size_t n = (size_t)width * height * depth;
  if (fwrite(data, sizeof(float), n, file.f) != n) return ws_fail(-13, std::string("short write to ") + path);
  return 0;
}

int b200ws_graph_write(const char* path, int width, int height, const int32_t* data, size_t count) {
  if (width <= 0 || height <= 0) return ws_fail(-11, "non-positive graph size");
  File file(path, "wb");
  if (!file.f) return ws_fail(-10, std::string("cannot open ") + path);
  fprintf(file.f, "%d&%d&%d&", width, height, 1);
  if (count && fwrite(data, sizeof(int32_t), count, file.f) != count) return ws_fail(-13, std::string("short write to ") + path);
  return 0;
}

int b200ws_graph_read(const char* path, int* width, int* height, int32_t* data, size_t capacity, size_t* count) {
  File file(path, "rb");
  if (!file.f) return ws_fail(-10, std::string("cannot open ") + path);
  long long v[3];
  if (!read_header(file.f, v) || v[0] <= 0 || v[1] <= 0 || v[2] <= 0) return ws_fail(-11, std::string("malformed graph header in ") + path);
  const long pos = ftell(file.f);
  fseek(file.f, 0, SEEK_END);
  const size_t n = (size_t)(ftell(file.f) - pos) / sizeof(int32_t);
  *width = (int)v[0]; *height = (int)v[1]; *count = n;
  if (capacity == 0) return 0;
  if (capacity < n) return ws_fail(-12, "buffer too small for the graph");
  fseek(file.f, pos, SEEK_SET);
  if (n && fread(data, sizeof(int32_t), n, file.f) != n) return ws_fail(-13, std::string("truncated graph file ") + path);
  return 0;
}

int b200ws_graph_build_map(int width, int height, const int32_t* data, size_t count, int32_t* map) {
  for (size_t i = 0; i < (size_t)width * height; ++i) map[i] = -1;
  for (size_t i = 0; i < count;) {
    if (i + 2 >= count) return ws_fail(-14, "corrupt consistency graph: insufficient data");
    const int col = data[i], row = data[i + 1], n = data[i + 2];
    if (n < 0) return ws_fail(-14, "corrupt consistency graph: negative image count");
    if (col < 0 || col >= width || row < 0 || row >= height) return ws_fail(-14, "corrupt consistency graph: pixel outside the map");
    if (i + 3 + (size_t)n > count) return ws_fail(-14, "corrupt consistency graph: record runs past the end");
    if (n > 0) map[(size_t)row * width + col] = (int32_t)(i + 2);
    i += 3 + (size_t)n;
  }
  return 0;
}

int b200ws_compute_depth_ranges(const b200ws_model* m, float* ranges) {
  if (int rc = check_model(m)) return rc;
  std::vector<std::vector<float>> depths(m->num_images);
  for (int64_t p = 0; p < m->num_points; ++p) {
    const float x = m->xyz[3 * p], y = m->xyz[3 * p + 1], z = m->xyz[3 * p + 2];
    for (int64_t j = m->track_offset[p]; j < m->track_offset[p + 1]; ++j) {
      const int i = m->track[j];
      const float* R = m->R + 9 * (size_t)i;
      const float depth = (R[6] * x + R[7] * y + R[8] * z) + m->T[3 * (size_t)i + 2];   // third row of [R | T], fp32
      if (depth > 0) depths[i].push_back(depth);
    }
  }
  for (int i = 0; i < m->num_images; ++i) {
    std::vector<float>& d = depths[i];
    if (d.empty()) { ranges[2 * i] = -1.0f; ranges[2 * i + 1] = -1.0f; continue; }
    std::sort(d.begin(), d.end());
    // indices size * 0.01f / size * 0.99f evaluated in fp32 and truncated, as in the reference
    const size_t lo = (size_t)((float)d.size() * 0.01f), hi = (size_t)((float)d.size() * 0.99f);
    ranges[2 * i] = d[std::min(lo, d.size() - 1)] * (1.0f - 0.25f);
    ranges[2 * i + 1] = d[std::min(hi, d.size() - 1)] * (1.0f + 0.25f);
  }
  return 0;
}

int b200ws_compute_shared_points(const b200ws_model* m, int32_t* counts) {
  if (int rc = check_model(m)) return rc;
  std::vector<int32_t> c;
  pair_statistics(m, &c, nullptr);
  memcpy(counts, c.data(), sizeof(int32_t) * c.size());
  return 0;
}

int b200ws_compute_triangulation_angles(const b200ws_model* m, float percentile_value, float* angles) {
  if (int rc = check_model(m)) return rc;
  if (!(percentile_value >= 0.0f && percentile_value <= 100.0f)) return ws_fail(-3, "percentile outside [0, 100]");
  std::vector<float> a;
  tri_percentile_dense(m, percentile_value, &a, nullptr);
  memcpy(angles, a.data(), sizeof(float) * a.size());
  return 0;
}

int b200ws_max_overlapping_images(const b200ws_model* m, int max_num, double min_triangulation_angle_deg, int32_t* out,
                                  int32_t* out_count) {
  if (int rc = check_model(m)) return rc;
  if (max_num < 0) return ws_fail(-3, "negative image count");
  const int n = m->num_images;
  std::vector<float> tri;
  std::vector<int32_t> counts;
  tri_percentile_dense(m, 75.0f, &tri, &counts);
  const float min_rad = (float)(min_triangulation_angle_deg * 0.0174532925199432954743716805978692718781530857086181640625);
  for (int i = 0; i < n; ++i) {
    const std::vector<int> r = ranked_sources(i, n, counts, tri, min_rad, (size_t)max_num);
    for (int k = 0; k < max_num; ++k) out[(size_t)i * max_num + k] = k < (int)r.size() ? r[k] : -1;
    if (out_count) out_count[i] = (int32_t)r.size();
  }
  return 0;
}

int b200ws_read_problems(const char* config_text, const b200ws_model* m, const char* const* image_names,
                         double min_triangulation_angle_deg, int32_t* ref_idx, int64_t* src_offset, int32_t* src_idx,
                         size_t cap_problems, size_t cap_src, size_t* num_problems, size_t* num_src) {
  if (int rc = check_model(m)) return rc;
  if (!config_text || !image_names || !num_problems || !num_src) return ws_fail(-1, "null argument");
  const int n = m->num_images;
  std::unordered_map<std::string, int> name_to_idx;
  for (int i = 0; i < n; ++i) name_to_idx.emplace(image_names[i], i);
  auto lookup = [&](const std::string& name, int* idx) {
    auto it = name_to_idx.find(name);
    if (it == name_to_idx.end()) { ws_fail(-20, "Image with name `" + name + "` does not exist"); return false; }
    *idx = it->second;
    return true;
  };
  // pass 1: pairs of (reference line, source line)
  std::vector<std::pair<std::string, std::vector<std::string>>> configs;
  {
    std::string ref;
    const char* p = config_text;
    while (*p) {
      const char* e = strchr(p, '\n');
      std::string line = trim(e ? std::string(p, e) : std::string(p));
      p = e ? e + 1 : p + strlen(p);
      if (line.empty() || line[0] == '#') continue;
      if (ref.empty()) { ref = line; continue; }
      int dummy;
      if (!lookup(ref, &dummy)) return -20;
      std::vector<std::string> items;   // CSVToVector<std::string> (util/misc.h:108-139): split at ',' or ';', trim, drop empty items
      size_t a = 0;
      while (a <= line.size()) {
        const size_t b = line.find_first_of(",;", a);
        const std::string item = trim(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (!item.empty()) items.push_back(item);
        if (b == std::string::npos) break;
        a = b + 1;
      }
      configs.emplace_back(ref, items);
      ref.clear();
    }
  }
  std::vector<float> tri;
  std::vector<int32_t> counts;
  const float min_rad = (float)(min_triangulation_angle_deg * 0.0174532925199432954743716805978692718781530857086181640625);
  std::vector<int32_t> refs;
  std::vector<std::vector<int>> srcs;
  for (const auto& cfg : configs) {
    int ref;
    if (!lookup(cfg.first, &ref)) return -20;
    std::vector<int> src;
    const std::vector<std::string>& it = cfg.second;
    if (it.size() == 1 && it[0] == "__all__") {
      for (int i = 0; i < n; ++i) if (i != ref) src.push_back(i);
    } else if (it.size() == 2 && it[0] == "__auto__") {
      if (tri.empty() && n > 0) tri_percentile_dense(m, 75.0f, &tri, &counts);
      char* end = nullptr;
      const long long max_num = strtoll(it[1].c_str(), &end, 10);
      if (end == it[1].c_str() || max_num < 0) return ws_fail(-21, "malformed __auto__ source specification: " + it[1]);
      src = ranked_sources(ref, n, counts, tri, min_rad, (size_t)max_num);
    } else {
      for (const std::string& name : it) { int idx; if (!lookup(name, &idx)) return -20; src.push_back(idx); }
    }
    if (src.empty()) continue;   // "Ignoring reference image ..., because it has no source images."
    refs.push_back(ref);
    srcs.push_back(src);
  }
  size_t total = 0;
  for (const auto& s : srcs) total += s.size();
  *num_problems = refs.size(); *num_src = total;
  if (cap_problems == 0) return 0;
  if (cap_problems < refs.size() || cap_src < total) return ws_fail(-12, "buffers too small for the problem list");
  int64_t off = 0;
  for (size_t k = 0; k < refs.size(); ++k) {
    ref_idx[k] = refs[k]; src_offset[k] = off;
    for (int v : srcs[k]) src_idx[off++] = v;
  }
  src_offset[refs.size()] = off;
  return 0;
}

}  // extern "C"
