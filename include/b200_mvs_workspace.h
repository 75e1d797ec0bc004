/*
 * b200_mvs_workspace.h — C-ABI of the host-side pieces either side of the PatchMatch sweep:
 * the on-disk map formats, the sparse-model statistics that pick source images and depth ranges, and the
 * `patch-match.cfg` problem list.  Plain pointers and sizes; 0 == success, negative == error
 * (b200ws_last_error() has the text).  No CUDA in these entry points.
 *
 * Reference interfaces replaced:
 *   mvs::Mat<float>::Read / Write                    src/colmap/mvs/mat.cc:41-66           (depth / normal / sel-prob maps)
 *   mvs::ConsistencyGraph::Read / Write / map        src/colmap/mvs/consistency_graph.cc:66-139
 *   mvs::Model::ComputeDepthRanges                   src/colmap/mvs/model.cc:178-218
 *   mvs::Model::ComputeSharedPoints                  src/colmap/mvs/model.cc:220-235
 *   mvs::Model::ComputeTriangulationAngles           src/colmap/mvs/model.cc:237-283
 *   mvs::Model::GetMaxOverlappingImages              src/colmap/mvs/model.cc:120-171
 *   PatchMatchController::ReadProblems               src/colmap/mvs/patch_match.cc:240-372 (patch-match.cfg)
 */
#ifndef B200_MVS_WORKSPACE_H_
#define B200_MVS_WORKSPACE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- map files: ASCII header "W&H&D&" followed by W*H*D little-endian float32, slice-major then row-major ---- */
int b200ws_mat_read_header(const char* path, int* width, int* height, int* depth);
int b200ws_mat_read(const char* path, float* data, size_t capacity /* floats */);
int b200ws_mat_write(const char* path, const float* data, int width, int height, int depth);

/* ---- consistency graph: header "W&H&1&" then int32 records [col, row, n, idx_1 .. idx_n] ---- */
int b200ws_graph_write(const char* path, int width, int height, const int32_t* data, size_t count);
/* two-call pattern: capacity 0 returns the record count in *count */
int b200ws_graph_read(const char* path, int* width, int* height, int32_t* data, size_t capacity, size_t* count);
/* ConsistencyGraph::InitializeMap: map[row*width + col] = offset of the pixel's `n` inside data, or -1; validates */
int b200ws_graph_build_map(int width, int height, const int32_t* data, size_t count, int32_t* map);

/* ---- sparse model (mvs::Model): images = R (9, row-major) + T (3) float32; points = xyz float32 + image-index tracks ---- */
typedef struct b200ws_model {
  int num_images;
  const float* R;               /* [9 * num_images] */
  const float* T;               /* [3 * num_images] */
  int64_t num_points;
  const float* xyz;             /* [3 * num_points] */
  const int64_t* track_offset;  /* [num_points + 1] into track */
  const int32_t* track;         /* image indices */
} b200ws_model;

/* ranges[2*i] / ranges[2*i+1] = depth_min / depth_max of image i (1st / 99th percentile of the positive depths of its
 * points, stretched by 25 %), -1 / -1 without points */
int b200ws_compute_depth_ranges(const b200ws_model* m, float* ranges);
/* dense [num_images x num_images] count of shared points */
int b200ws_compute_shared_points(const b200ws_model* m, int32_t* counts);
/* dense [num_images x num_images] percentile (0..100) of the pairwise triangulation angles in radians, -1 where the
 * two images share no point */
int b200ws_compute_triangulation_angles(const b200ws_model* m, float percentile, float* angles);
/* for every image the (at most max_num) images with the most shared points among those whose 75th-percentile
 * triangulation angle is >= min_triangulation_angle_deg; out is [num_images x max_num], -1 padded.  Equal counts are
 * ordered by ascending image index (the reference's std::partial_sort leaves that order unspecified). */
int b200ws_max_overlapping_images(const b200ws_model* m, int max_num, double min_triangulation_angle_deg, int32_t* out,
                                  int32_t* out_count);

/* ---- patch-match.cfg: alternating lines "reference image name" / "source spec" ('#' comments, blank lines skipped);
 * source spec = "__all__" | "__auto__, N" | comma-separated image names.  Problems without source images are dropped.
 * Outputs: ref_idx[num_problems], src_offset[num_problems + 1], src_idx[...]; two-call pattern: with cap_problems == 0
 * only *num_problems and *num_src are written. */
int b200ws_read_problems(const char* config_text, const b200ws_model* m, const char* const* image_names,
                         double min_triangulation_angle_deg, int32_t* ref_idx, int64_t* src_offset, int32_t* src_idx,
                         size_t cap_problems, size_t cap_src, size_t* num_problems, size_t* num_src);

const char* b200ws_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_MVS_WORKSPACE_H_ */
