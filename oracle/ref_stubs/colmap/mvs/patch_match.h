// Stub of colmap/mvs/patch_match.h: only PatchMatchOptions (the real header) and PatchMatch::Problem
// (patch_match.h:57-75), without the controller (threading / workspace dependencies).
#pragma once
#include "colmap/mvs/depth_map.h"
#include "colmap/mvs/image.h"
#include "colmap/mvs/normal_map.h"
#include "colmap/mvs/patch_match_options.h"
#include <vector>
namespace colmap {
namespace mvs {
const static size_t kMaxPatchMatchWindowRadius = 32;
class PatchMatch {
 public:
  struct Problem {
    int ref_image_idx = -1;
    std::vector<int> src_image_idxs;
    std::vector<Image>* images = nullptr;
    std::vector<DepthMap>* depth_maps = nullptr;
    std::vector<NormalMap>* normal_maps = nullptr;
  };
};
}  // namespace mvs
}  // namespace colmap
