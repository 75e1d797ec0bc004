// Stub of colmap/sensor/bitmap.h (OpenImageIO is not installed): an 8-bit grey row-major buffer with the
// accessors the PatchMatch CUDA code uses.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
namespace colmap {
class Bitmap {
 public:
  Bitmap() {}
  Bitmap(int width, int height, bool as_rgb, bool = false) : width_(width), height_(height), channels_(as_rgb ? 3 : 1), data_((size_t)width * height * (as_rgb ? 3 : 1)) {}
  int Width() const { return width_; }
  int Height() const { return height_; }
  int Channels() const { return channels_; }
  size_t NumBytes() const { return data_.size(); }
  bool IsGrey() const { return channels_ == 1; }
  bool IsRGB() const { return channels_ == 3; }
  bool IsEmpty() const { return data_.empty(); }
  std::vector<uint8_t>& RowMajorData() { return data_; }
  const std::vector<uint8_t>& RowMajorData() const { return data_; }
 private:
  int width_ = 0, height_ = 0, channels_ = 1;
  std::vector<uint8_t> data_;
};
}  // namespace colmap
