/*
 * b200_mvs_fusion.h — C-ABI of the stereo fusion step that consumes the sweep's outputs.
 * Host code (the reference's StereoFusion is a CPU traversal with a visited mask; its order is part of the result).
 *
 * Reference interfaces replaced:
 *   StereoFusionOptions                       src/colmap/mvs/fusion.h:47-95, Check: fusion.cc:97-107
 *   StereoFusion::Run / Fuse                  src/colmap/mvs/fusion.cc:131-336 (image order, row scan), :367-545 (traversal)
 *   internal::FindNextImage                   src/colmap/mvs/fusion.cc:48-72
 *   WritePointsVisibility / ReadPointsVisibility   src/colmap/mvs/fusion.cc:547-588
 * Semantics: the single-threaded schedule of the reference (num_threads = 1: rows top to bottom, columns left to right,
 * images in FindNextImage order), which is its only deterministic one.  Visibility lists come back sorted ascending
 * (the reference iterates a hash set).  0 == success, negative == error (b200fuse_last_error()).
 */
#ifndef B200_MVS_FUSION_H_
#define B200_MVS_FUSION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200fuse_options {
  int min_num_pixels;        /* 5 */
  int max_num_pixels;        /* 10000 */
  int max_traversal_depth;   /* 100 */
  double max_reproj_error;   /* 2.0 px */
  double max_depth_error;    /* 0.01 relative */
  double max_normal_error;   /* 10 degrees */
  float bbox_min[3];         /* -FLT_MAX */
  float bbox_max[3];         /* +FLT_MAX */
} b200fuse_options;

typedef struct b200fuse_image {
  int used;                  /* listed in fusion.cfg and all inputs present */
  int image_width, image_height;   /* size the calibration K refers to */
  const float* K;            /* 9, row-major */
  const float* R;            /* 9 */
  const float* T;            /* 3 */
  int map_width, map_height; /* depth / normal map size */
  const float* depth;        /* [map_height * map_width] */
  const float* normal;       /* [3 * map_height * map_width], slice-major (camera frame) */
  int bitmap_width, bitmap_height;
  const uint8_t* rgb;        /* [bitmap_height * bitmap_width * 3] */
  const uint8_t* mask;       /* [map_height * map_width] non-zero = pre-masked pixel, or NULL */
} b200fuse_image;

typedef struct b200fuse_result* b200fuse_result_t;

void b200fuse_options_init(b200fuse_options* o);
int b200fuse_options_check(const b200fuse_options* o);   /* 1 == valid */

/* overlap: [num_images x max_overlap] image indices, -1 padded (Model::GetMaxOverlappingImages(check_num_images, 0)) */
int b200fuse_run(const b200fuse_options* o, int num_images, const b200fuse_image* images, const int32_t* overlap,
                 int max_overlap, b200fuse_result_t* out);
int64_t b200fuse_num_points(b200fuse_result_t r);
int64_t b200fuse_num_visibility(b200fuse_result_t r);   /* total length of all visibility lists */
/* xyz [3n], normal [3n], rgb [3n], vis_offset [n + 1], vis [num_visibility] */
int b200fuse_get(b200fuse_result_t r, float* xyz, float* normal, uint8_t* rgb, int64_t* vis_offset, int32_t* vis);
void b200fuse_free(b200fuse_result_t r);

/* <num_points u64> then per point <n u32><image_idx u32 ...> */
int b200fuse_write_visibility(const char* path, int64_t num_points, const int64_t* vis_offset, const int32_t* vis);
/* two-call pattern: vis == NULL returns the total length in *num_visibility; fails if the file's count != num_points */
int b200fuse_read_visibility(const char* path, int64_t num_points, int64_t* vis_offset, int32_t* vis, int64_t* num_visibility);

const char* b200fuse_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_MVS_FUSION_H_ */
