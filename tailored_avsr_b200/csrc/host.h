// Host-side helpers shared by the C-ABI entry points: thread-local error string, CUDA error
// checking, TMA tensor-map encoding (driver entry point resolved at run time so the library links
// against cudart only) with a small cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/tavsr.h"

namespace tavsr {

extern thread_local char g_last_error[512];
extern int g_debug[16];

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

#define TAVSR_CUDA_OK(expr)                                                               \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::tavsr::set_error(TAVSR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                                cudaGetErrorString(_e), __FILE__, __LINE__);              \
  } while (0)

#define TAVSR_REQUIRE(cond, ...)                                            \
  do {                                                                      \
    if (!(cond)) return ::tavsr::set_error(TAVSR_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Encode (or fetch from cache) a 2-D row-major tensor map:
//   inner dim = `cols` elements (contiguous), outer dim = `rows`, row pitch `ld` elements.
//   box = {box_cols, box_rows}, 128-byte swizzle (box_cols * elem_bytes must be 128).
// Returns 0 on success.
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, bool is_bf16, uint64_t rows,
                 uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                 bool is_load = true);

int num_sms();

// Function attributes (MaxDynamicSharedMemorySize) are per device: `mask` keeps one bit per device
// ordinal, the caller sets the attribute when this returns true.
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// Same for attributes whose value grows with the problem (dynamic shared memory sized by T / V / D):
// returns true when device `dev`'s recorded size is below `bytes` (and records it).
struct PerDeviceMax {
  int v[64] = {0};
  bool raise(int bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (v[dev & 63] >= bytes) return false;
    v[dev & 63] = bytes;
    return true;
  }
};

// Uniform launch path: optional thread-block cluster, optional programmatic dependent launch
// (g_debug[7] == 0 enables PDL; set it to 1 to fall back to plain stream-ordered launches).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  int n = 0;
  if (cluster > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster;
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_debug[7] == 0) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace tavsr
