// pm_workspace.cu — workspace-level drop-in: PatchMatchController::Run behind b200pm_run_workspace (host code).
//
// Reference behaviour: src/colmap/mvs/patch_match.cc:156-536 (controller), mvs/model.cc:56-100 (sparse model ->
// mvs::Model), mvs/workspace.cc (map file names).  Everything numeric happens in b200pm_* (patch_match.cu) and
// b200ws_* (mvs_workspace.cu); this file is the scheduling and the I/O either side of them.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200_mvs_workspace.h"
#include "../../include/b200_patch_match.h"

extern "C" void b200pm_internal_set_error(const char* msg);

namespace {

int fail(int code, const std::string& m) { b200pm_internal_set_error(m.c_str()); return code; }

struct WsImage {
  std::string name, path;
  int width = 0, height = 0;
  float K[9], R[9], T[3];
};
struct WsModel {
  std::vector<WsImage> images;
  std::vector<float> R, T, xyz;
  std::vector<int64_t> track_offset;
  std::vector<int32_t> track;
};

bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
void make_dirs(const std::string& p) {
  for (size_t i = 1; i <= p.size(); ++i)
    if (i == p.size() || p[i] == '/') mkdir(p.substr(0, i).c_str(), 0777);
}
template <typename T> bool rd(FILE* f, T* v, size_t n = 1) { return fread(v, sizeof(T), n, f) == n; }

// number of parameters of a CameraModelId (sensor/models.h:90-109) and where focal length(s) / principal point sit
int model_num_params(int id) {
  static const int n[] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};
  return (id >= 0 && id <= 10) ? n[id] : -1;
}

// mvs::Model::ReadFromCOLMAP (model.cc:56-100) from the legacy binary trio (scene/reconstruction_io_binary.cc:107-291):
// images in ascending image id, K = Camera::CalibrationMatrix, R / T of cam_from_world in float
int read_colmap_model(const std::string& ws, WsModel* M) {
  struct Cam { int model; uint64_t w, h; double p[12]; };
  std::map<uint32_t, Cam> cams;
  {
    FILE* f = fopen((ws + "/sparse/cameras.bin").c_str(), "rb");
    if (!f) return fail(-20, "cannot open " + ws + "/sparse/cameras.bin");
    uint64_t n = 0; rd(f, &n);
    for (uint64_t i = 0; i < n; ++i) {
      uint32_t id; int32_t model; Cam c;
      if (!rd(f, &id) || !rd(f, &model) || !rd(f, &c.w) || !rd(f, &c.h)) { fclose(f); return fail(-20, "cameras.bin: truncated"); }
      const int np = model_num_params(model);
      if (np < 0 || !rd(f, c.p, np)) { fclose(f); return fail(-20, "cameras.bin: unknown camera model"); }
      c.model = model; cams[id] = c;
    }
    fclose(f);
  }
  struct Img { double q[4], t[3]; uint32_t cam; std::string name; };
  std::map<uint32_t, Img> imgs;
  {
    FILE* f = fopen((ws + "/sparse/images.bin").c_str(), "rb");
    if (!f) return fail(-20, "cannot open " + ws + "/sparse/images.bin");
    uint64_t n = 0; rd(f, &n);
    for (uint64_t i = 0; i < n; ++i) {
      uint32_t id; Img im;
      if (!rd(f, &id) || !rd(f, im.q, 4) || !rd(f, im.t, 3) || !rd(f, &im.cam)) { fclose(f); return fail(-20, "images.bin: truncated"); }
      for (int ch; (ch = fgetc(f)) > 0;) im.name.push_back((char)ch);
      uint64_t m = 0;
      if (!rd(f, &m) || m > (1ULL << 32)) { fclose(f); return fail(-20, "images.bin: truncated"); }   // (a corrupt count must not wrap the seek)
      if (fseek(f, (long)(24 * m), SEEK_CUR) != 0) { fclose(f); return fail(-20, "images.bin: truncated"); }
      imgs[id] = im;
    }
    fclose(f);
  }
  std::map<uint32_t, int> id_to_idx;
  for (const auto& kv : imgs) {
    const Img& im = kv.second;
    const auto ci = cams.find(im.cam);
    if (ci == cams.end()) return fail(-20, "images.bin references an unknown camera");
    const Cam& c = ci->second;
    WsImage w; w.name = im.name; w.path = ws + "/images/" + im.name; w.width = (int)c.w; w.height = (int)c.h;
    const bool single = c.model == 0 || c.model == 2 || c.model == 3 || c.model == 8 || c.model == 9;
    const double fx = c.p[0], fy = single ? c.p[0] : c.p[1], cx = single ? c.p[1] : c.p[2], cy = single ? c.p[2] : c.p[3];
    const float K[9] = {(float)fx, 0, (float)cx, 0, (float)fy, (float)cy, 0, 0, 1};
    memcpy(w.K, K, sizeof(K));
    double qw = im.q[0], qx = im.q[1], qy = im.q[2], qz = im.q[3];
    const double nq = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= nq; qx /= nq; qy /= nq; qz /= nq;
    const double Rd[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy),
                          2 * (qx * qy + qw * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx),
                          2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)};
    for (int k = 0; k < 9; ++k) w.R[k] = (float)Rd[k];
    for (int k = 0; k < 3; ++k) w.T[k] = (float)im.t[k];
    id_to_idx[kv.first] = (int)M->images.size();
    M->images.push_back(w);
  }
  {
    FILE* f = fopen((ws + "/sparse/points3D.bin").c_str(), "rb");
    if (!f) return fail(-20, "cannot open " + ws + "/sparse/points3D.bin");
    uint64_t n = 0; rd(f, &n);
    M->track_offset.push_back(0);
    for (uint64_t i = 0; i < n; ++i) {
      uint64_t pid; double xyz[3]; uint8_t rgb[3]; double err; uint64_t tl;
      if (!rd(f, &pid) || !rd(f, xyz, 3) || !rd(f, rgb, 3) || !rd(f, &err) || !rd(f, &tl)) { fclose(f); return fail(-20, "points3D.bin: truncated"); }
      for (int k = 0; k < 3; ++k) M->xyz.push_back((float)xyz[k]);
      for (uint64_t t = 0; t < tl; ++t) {
        uint32_t el[2];
        if (!rd(f, el, 2)) { fclose(f); return fail(-20, "points3D.bin: truncated"); }
        const auto it = id_to_idx.find(el[0]);
        if (it == id_to_idx.end()) { fclose(f); return fail(-20, "points3D.bin references an unknown image"); }
        M->track.push_back(it->second);
      }
      M->track_offset.push_back((int64_t)M->track.size());
    }
    fclose(f);
  }
  for (const WsImage& w : M->images) { M->R.insert(M->R.end(), w.R, w.R + 9); M->T.insert(M->T.end(), w.T, w.T + 3); }
  return 0;
}

// built-in bitmap reader: binary PGM (P5), 8 bit
int load_pgm(void*, const char* path, int* w, int* h, uint8_t** data) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  char magic[3] = {0, 0, 0};
  int maxv = 0;
  auto skip = [&]() { int c; while ((c = fgetc(f)) != EOF) { if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} } else if (c > ' ') { ungetc(c, f); break; } } };
  if (fread(magic, 1, 2, f) != 2 || magic[0] != 'P' || magic[1] != '5') { fclose(f); return -1; }
  skip(); if (fscanf(f, "%d", w) != 1) { fclose(f); return -1; }
  skip(); if (fscanf(f, "%d", h) != 1) { fclose(f); return -1; }
  skip(); if (fscanf(f, "%d", &maxv) != 1 || maxv != 255) { fclose(f); return -1; }
  fgetc(f);
  if (*w <= 0 || *h <= 0 || *w > (1 << 16) || *h > (1 << 16)) { fclose(f); return -1; }   // a corrupt header must not size the allocation
  *data = (uint8_t*)malloc((size_t)*w * *h);
  if (!*data) { fclose(f); return -1; }
  const bool ok = fread(*data, 1, (size_t)*w * *h, f) == (size_t)*w * *h;
  fclose(f);
  if (!ok) { free(*data); *data = nullptr; return -1; }
  return 0;
}

struct Bitmap { int w = 0, h = 0; std::vector<uint8_t> px; bool tried = false, ok = false; };

struct Controller {
  const b200pm_options* opt;
  const b200pm_workspace* ws;
  std::string root, stereo;
  WsModel M;
  std::vector<float> ranges;
  std::vector<int32_t> ref_idx, src_idx;
  std::vector<int64_t> src_off;
  std::vector<Bitmap> bitmaps;
  std::mutex io_mutex;   // the reference serialises workspace access too (workspace_mutex_, patch_match.cc:463)
  std::atomic<int> processed{0};
  std::mutex err_mutex;
  std::string first_error;

  std::string out_path(const char* kind, int image, const char* type) const {
    return root + "/" + stereo + "/" + kind + "/" + M.images[image].name + "." + type + ".bin";
  }
  const Bitmap* bitmap(int idx) {
    Bitmap& b = bitmaps[idx];
    if (!b.tried) {
      b.tried = true;
      uint8_t* d = nullptr;
      b200pm_load_gray_fn fn = ws->load_gray ? ws->load_gray : load_pgm;
      if (fn(ws->load_gray_user, M.images[idx].path.c_str(), &b.w, &b.h, &d) == 0 && d) {
        b.px.assign(d, d + (size_t)b.w * b.h); free(d); b.ok = true;
      }
    }
    return b.ok ? &b : nullptr;
  }
  void note(const std::string& m) { std::lock_guard<std::mutex> l(err_mutex); if (first_error.empty()) first_error = m; }

  // PatchMatchController::ProcessProblem (:385-535)
  void process(int k, bool geom, bool filter, int gpu) {
    const int ref = ref_idx[k];
    const char* type = geom ? "geometric" : "photometric";
    const std::string dp = out_path("depth_maps", ref, type), np = out_path("normal_maps", ref, type), gp = out_path("consistency_graphs", ref, type);
    if (file_exists(dp) && file_exists(np) && (!ws->write_consistency_graph || file_exists(gp))) return;
    b200pm_options o = *opt;
    o.geom_consistency = geom ? 1 : 0; o.filter = filter ? 1 : 0; o.gpu_index = gpu;
    if (o.depth_min < 0 || o.depth_max < 0) {
      o.depth_min = ranges[2 * ref]; o.depth_max = ranges[2 * ref + 1];
      if (!(o.depth_min > 0 && o.depth_max > 0)) { note("You must manually set the minimum and maximum depth, since no sparse model is provided in the workspace."); return; }
    }
    if (o.sigma_spatial <= 0) o.sigma_spatial = o.window_radius;
    std::vector<int> srcs;   // used_image_idxs is a set: duplicates collapse, the reference image is no source of itself
    for (int64_t s = src_off[k]; s < src_off[k + 1]; ++s)
      if (src_idx[s] != ref && std::find(srcs.begin(), srcs.end(), src_idx[s]) == srcs.end()) srcs.push_back(src_idx[s]);
    o.filter_min_num_consistent = std::min<int>((int)srcs.size(), o.filter_min_num_consistent);
    const int n = (int)srcs.size();
    b200pm_problem P; memset(&P, 0, sizeof(P));
    std::vector<int> sw(n), sh(n);
    std::vector<const uint8_t*> sg(n);
    std::vector<float> sK(9 * (size_t)n), sR(9 * (size_t)n), sT(3 * (size_t)n);
    std::vector<std::vector<float>> sdepth(n);
    std::vector<const float*> sdp(n);
    std::vector<float> rdepth, rnormal;
    {
      std::lock_guard<std::mutex> l(io_mutex);
      auto scaled_K = [&](int idx, const Bitmap* b, float* K) {   // Image::Rescale (mvs/image.cc:66-95)
        memcpy(K, M.images[idx].K, 9 * sizeof(float));
        if (b->w != M.images[idx].width || b->h != M.images[idx].height) {
          const float sx = (float)b->w / M.images[idx].width, sy = (float)b->h / M.images[idx].height;
          K[0] *= sx; K[2] *= sx; K[4] *= sy; K[5] *= sy;
        }
      };
      const Bitmap* rb = bitmap(ref);
      if (!rb) { note("cannot read image " + M.images[ref].path); return; }
      P.ref_width = rb->w; P.ref_height = rb->h; P.ref_gray = rb->px.data();
      scaled_K(ref, rb, P.ref_K); memcpy(P.ref_R, M.images[ref].R, sizeof(P.ref_R)); memcpy(P.ref_T, M.images[ref].T, sizeof(P.ref_T));
      for (int i = 0; i < n; ++i) {
        const Bitmap* b = bitmap(srcs[i]);
        if (!b) { note("cannot read image " + M.images[srcs[i]].path); return; }
        sw[i] = b->w; sh[i] = b->h; sg[i] = b->px.data();
        scaled_K(srcs[i], b, sK.data() + 9 * i);
        memcpy(sR.data() + 9 * i, M.images[srcs[i]].R, 9 * sizeof(float)); memcpy(sT.data() + 3 * i, M.images[srcs[i]].T, 3 * sizeof(float));
      }
      if (geom) {   // the photometric maps of the reference and of every source image (workspace.cc:208-240)
        auto read_map = [&](const std::string& path, int want_depth, std::vector<float>& out) {
          int w = 0, h = 0, d = 0;
          if (b200ws_mat_read_header(path.c_str(), &w, &h, &d) != 0 || d != want_depth) return false;
          out.resize((size_t)w * h * d);
          return b200ws_mat_read(path.c_str(), out.data(), out.size()) == 0;
        };
        if (!read_map(out_path("depth_maps", ref, "photometric"), 1, rdepth) || !read_map(out_path("normal_maps", ref, "photometric"), 3, rnormal)) { note("missing photometric maps of " + M.images[ref].name); return; }
        for (int i = 0; i < n; ++i) {
          if (!read_map(out_path("depth_maps", srcs[i], "photometric"), 1, sdepth[i])) { note("missing photometric depth map of " + M.images[srcs[i]].name); return; }
          sdp[i] = sdepth[i].data();
        }
      }
    }
    P.num_src = n; P.src_width = sw.data(); P.src_height = sh.data(); P.src_gray = sg.data();
    P.src_K = sK.data(); P.src_R = sR.data(); P.src_T = sT.data(); P.src_image_idxs = srcs.data();
    if (geom) { P.src_depth = sdp.data(); P.ref_depth_init = rdepth.data(); P.ref_normal_init = rnormal.data(); }
    b200pm_handle h = nullptr;
    if (b200pm_create(&o, &P, &h) != 0 || b200pm_run(h) != 0) { note(std::string("PatchMatch failed for ") + M.images[ref].name + ": " + b200pm_last_error()); if (h) b200pm_destroy(h); return; }
    const size_t npx = (size_t)P.ref_width * P.ref_height;
    std::vector<float> depth(npx), normal(3 * npx);
    int* graph = nullptr; size_t ngraph = 0;
    int rc = b200pm_get_depth(h, depth.data());
    if (rc == 0) rc = b200pm_get_normal(h, normal.data());
    if (rc == 0 && ws->write_consistency_graph && filter) rc = b200pm_get_consistency(h, &graph, &ngraph);
    b200pm_destroy(h);
    if (rc != 0) { note(std::string("reading results failed: ") + b200pm_last_error()); return; }
    for (const std::string& p : {dp, np, gp}) make_dirs(p.substr(0, p.rfind('/')));
    if (b200ws_mat_write(dp.c_str(), depth.data(), P.ref_width, P.ref_height, 1) != 0 ||
        b200ws_mat_write(np.c_str(), normal.data(), P.ref_width, P.ref_height, 3) != 0) { note(std::string("writing maps failed: ") + b200ws_last_error()); return; }
    if (ws->write_consistency_graph) b200ws_graph_write(gp.c_str(), P.ref_width, P.ref_height, graph, ngraph);
    if (graph) b200pm_free(graph);
    processed.fetch_add(1);
  }

  void run_phase(bool geom, bool filter, const std::vector<int>& mine, const std::vector<int>& gpus) {
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int gpu : gpus)
      pool.emplace_back([&, gpu]() { for (size_t i; (i = next.fetch_add(1)) < mine.size();) process(mine[i], geom, filter, gpu); });
    for (auto& t : pool) t.join();
  }
};

}  // namespace

extern "C" int b200pm_run_workspace(const b200pm_options* options, const b200pm_workspace* w, int* num_processed) {
  if (!options || !w || !w->workspace_path) return fail(-1, "null argument");
  Controller C;
  C.opt = options; C.ws = w; C.root = w->workspace_path; C.stereo = w->stereo_folder ? w->stereo_folder : "stereo";
  int rc = read_colmap_model(C.root, &C.M);
  if (rc != 0) return rc;
  const int NI = (int)C.M.images.size();
  b200ws_model m; memset(&m, 0, sizeof(m));
  m.num_images = NI; m.R = C.M.R.data(); m.T = C.M.T.data(); m.num_points = (int64_t)C.M.xyz.size() / 3; m.xyz = C.M.xyz.data();
  m.track_offset = C.M.track_offset.data(); m.track = C.M.track.data();
  C.ranges.resize(2 * (size_t)NI);
  if (b200ws_compute_depth_ranges(&m, C.ranges.data()) != 0) return fail(-21, std::string("depth ranges: ") + b200ws_last_error());
  {
    const std::string cfg_path = w->config_path ? w->config_path : C.root + "/" + C.stereo + "/patch-match.cfg";
    FILE* f = fopen(cfg_path.c_str(), "rb");
    if (!f) return fail(-22, "cannot open " + cfg_path);
    std::string text; char buf[4096]; size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    fclose(f);
    std::vector<const char*> names(NI);
    for (int i = 0; i < NI; ++i) names[i] = C.M.images[i].name.c_str();
    size_t np = 0, ns = 0;
    if (b200ws_read_problems(text.c_str(), &m, names.data(), options->min_triangulation_angle, nullptr, nullptr, nullptr, 0, 0, &np, &ns) != 0)
      return fail(-22, std::string("patch-match.cfg: ") + b200ws_last_error());
    C.ref_idx.resize(np); C.src_off.resize(np + 1); C.src_idx.resize(ns ? ns : 1);
    if (np && b200ws_read_problems(text.c_str(), &m, names.data(), options->min_triangulation_angle, C.ref_idx.data(), C.src_off.data(), C.src_idx.data(), np, ns, &np, &ns) != 0)
      return fail(-22, std::string("patch-match.cfg: ") + b200ws_last_error());
  }
  C.bitmaps.resize(NI);
  // this process' share of the problems: largest reference image x source count first onto the least loaded rank
  const int world = std::max(1, w->world_size), rank = std::min(std::max(0, w->rank), world - 1);
  std::vector<int> mine;
  {
    const int np = (int)C.ref_idx.size();
    std::vector<int> order(np);
    std::iota(order.begin(), order.end(), 0);
    auto cost = [&](int k) { return (double)C.M.images[C.ref_idx[k]].width * C.M.images[C.ref_idx[k]].height * (double)std::max<int64_t>(1, C.src_off[k + 1] - C.src_off[k]); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    std::vector<double> load(world, 0.0);
    for (int k : order) { const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin()); load[r] += cost(k); if (r == rank) mine.push_back(k); }
    std::sort(mine.begin(), mine.end());
  }
  std::vector<int> gpus(w->gpu_indices, w->gpu_indices + (w->gpu_indices ? w->num_gpu_indices : 0));
  if (gpus.empty()) gpus.push_back(options->gpu_index);
  const bool geom = options->geom_consistency != 0;
  if (geom && w->phase != 2) C.run_phase(false, false, mine, gpus);
  if (!(geom && w->phase == 1)) C.run_phase(geom, options->filter != 0, mine, gpus);
  if (num_processed) *num_processed = C.processed.load();
  if (!C.first_error.empty()) return fail(-23, C.first_error);
  return 0;
}
