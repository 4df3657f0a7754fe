// common.cuh -- context, error plumbing and small device helpers shared by all kernels.
// B200 (sm_100a) only; there is deliberately no CPU path in this library.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/strugatzki_b200.h"

namespace sgz {

inline std::string &err_slot() {
  static thread_local std::string s;
  return s;
}

inline void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err_slot() = buf;
}

#define SGZ_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::sgz::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return SGZ_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define SGZ_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::sgz::set_error(__VA_ARGS__);    \
      return SGZ_ERR_INVALID;           \
    }                                   \
  } while (0)

#define SGZ_TRY(expr)             \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ < 0) return rc__;    \
  } while (0)

// fullToFeat / featToFull, FeatureCorrelationImpl.scala:38-39
inline int full_to_feat(int64_t n, int step) { return (int)((n + (step >> 1)) / step); }
inline int64_t feat_to_full(int64_t i, int step) { return i * (int64_t)step; }

// java.lang.Float.compare total order (NaN greatest, -0.0 < +0.0)
__host__ __device__ inline int32_t jfloat_bits(float x) {
  if (x != x) return 0x7fc00000;
#ifdef __CUDA_ARCH__
  return __float_as_int(x);
#else
  int32_t b;
  memcpy(&b, &x, 4);
  return b;
#endif
}
__host__ __device__ inline int jfloat_compare(float x, float y) {
  if (x < y) return -1;
  if (x > y) return 1;
  int32_t a = jfloat_bits(x), b = jfloat_bits(y);
  return a == b ? 0 : (a < b ? -1 : 1);
}

// monotone map float -> uint32 (non-NaN); used for packed atomicMax keys
__host__ __device__ inline uint32_t float_order_key(float x) {
#ifdef __CUDA_ARCH__
  uint32_t b = (uint32_t)__float_as_int(x);
#else
  uint32_t b;
  memcpy(&b, &x, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float float_from_order_key(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __int_as_float((int)b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

}  // namespace sgz

// ---------------------------------------------------------------------------------------------
// opaque handle types
// ---------------------------------------------------------------------------------------------
struct sgz_ctx {
  // lifetime: databases hold a reference to their context and jobs to their database; a destroy call on a
  // handle that is still referenced only marks it, the last child releases it (hosts with garbage collection
  // destroy handles in arbitrary order)
  std::atomic<int> refs{0};
  bool zombie = false;
  int device = 0;
  int smCount = 0;
  size_t smemOptin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;        // total kernels launched on this context
  int64_t callLaunches0 = 0;   // snapshot at begin_call
  float lastMs = 0.f;
  int64_t lastLaunches = 0;

  int bind() const {
    SGZ_CUDA(cudaSetDevice(device));
    return SGZ_OK;
  }
  int begin_call() {
    callLaunches0 = launches;
    SGZ_CUDA(cudaEventRecord(ev0, stream));
    return SGZ_OK;
  }
  int end_call() {
    SGZ_CUDA(cudaEventRecord(ev1, stream));
    SGZ_CUDA(cudaEventSynchronize(ev1));
    SGZ_CUDA(cudaEventElapsedTime(&lastMs, ev0, ev1));
    lastLaunches = launches - callLaunches0;
    return SGZ_OK;
  }
};

#define SGZ_LAUNCH_CHECK(ctx)                                                               \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      ::sgz::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return SGZ_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// simple owning device buffer
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n && p) return SGZ_OK;
    release();
    if (count == 0) return SGZ_OK;
    cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      n = 0;
      ::sgz::set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? SGZ_ERR_NOMEM : SGZ_ERR_CUDA;
    }
    n = count;
    return SGZ_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};
