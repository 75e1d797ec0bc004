// device_cache.h — a small per-device caching allocator for the two C-ABI entry points.
// A workspace run creates and destroys hundreds of equally sized problems (one PatchMatch handle per reference image,
// one BA solve per mapper step); cudaMalloc / cudaFree (the latter synchronises the device) then cost more than the
// uploads.  Freed blocks are kept per (device, byte size) and handed back on the next request; the cache is capped and
// can be dropped with b200_release_cached_memory().
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

struct B200DeviceCache {
  std::mutex mu;
  std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;
  size_t cached_bytes = 0;
  static constexpr size_t kMaxCachedBytes = 16ull << 30;

  static B200DeviceCache& get() { static B200DeviceCache c; return c; }

  cudaError_t alloc(void** p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lock(mu);
      auto it = free_blocks.find({dev, bytes});
      if (it != free_blocks.end() && !it->second.empty()) {
        *p = it->second.back();
        it->second.pop_back();
        cached_bytes -= bytes;
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // make room and retry once
      release();
      cudaGetLastError();
      e = cudaMalloc(p, bytes);
    }
    return e;
  }
  void free(void* p, size_t bytes) {
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (cached_bytes + bytes > kMaxCachedBytes) { cudaFree(p); return; }
    free_blocks[{dev, bytes}].push_back(p);
    cached_bytes += bytes;
  }
  void release() {
    std::lock_guard<std::mutex> lock(mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : free_blocks) {
      cudaSetDevice(kv.first.first);
      for (void* p : kv.second) cudaFree(p);
      kv.second.clear();
    }
    cudaSetDevice(cur);
    cached_bytes = 0;
  }
};
