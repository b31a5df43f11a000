// Shared helpers for libgeobi (sm_100a).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "geobi.h"

namespace geobi {

void set_error(const char* fmt, ...);  // api.cu

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// Carves typed, 256-byte aligned arrays out of a caller-provided workspace.
struct Carver {
  char* base;
  size_t used;
  size_t cap;
  Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), used(0), cap(bytes) {}
  template <typename T>
  T* take(size_t n) {
    size_t off = align256(used);
    used = off + n * sizeof(T);
    return reinterpret_cast<T*>(base + off);
  }
  bool ok() const { return base != nullptr ? used <= cap : used == 0; }
};
// Same arithmetic without memory: for *_ws_bytes queries.
struct Sizer {
  size_t used = 0;
  template <typename T>
  void take(size_t n) { used = align256(used) + n * sizeof(T); }
  size_t total() const { return align256(used) + 256; }
};

#define GEOBI_CUDA_OK(expr)                                                            \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      (void)cudaGetLastError(); /* clear the non-sticky error so that later launches are not blamed */ \
      ::geobi::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GEOBI_ERR_CUDA;                                                           \
    }                                                                                  \
  } while (0)

#define GEOBI_LAUNCH_OK(name)                                                          \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ::geobi::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
      return GEOBI_ERR_CUDA;                                                           \
    }                                                                                  \
  } while (0)

#define GEOBI_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::geobi::set_error(__VA_ARGS__);      \
      return GEOBI_ERR_INVALID;             \
    }                                       \
  } while (0)

// ---- internal cross-file entry points (graph.cu)
int scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st);
size_t scan_ws_bytes(int64_t n);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace geobi

// The head projections P = X.U^T are computed in fp64 (feast_project_kernel) and stored as a double-float pair packed in the
// 8 bytes of a double: low word = hi = (float)P, high word = lo = (float)(P - hi)  (|P - hi - lo| <= 2^-48 |P|).
// p_diff(a, b) = fl(a - b) to within ~1.5 ulp of the DIFFERENCE on the fp32 pipe (3 FADD) - what the aggregation kernels
// need for the soft assignments - instead of a DADD + F2F on the fp64 pipe per head and edge.
#ifdef __CUDACC__
__device__ __forceinline__ double p_pack(double v) {
  const float hi = (float)v;
  const float lo = (float)(v - (double)hi);
  return __hiloint2double(__float_as_int(lo), __float_as_int(hi));
}
__device__ __forceinline__ float p_diff(double a, double b) {
  const float ah = __int_as_float(__double2loint(a)), al = __int_as_float(__double2hiint(a));
  const float bh = __int_as_float(__double2loint(b)), bl = __int_as_float(__double2hiint(b));
  return (ah - bh) + (al - bl);
}
#endif
