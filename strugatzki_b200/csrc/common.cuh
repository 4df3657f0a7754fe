// common.cuh -- context, error plumbing and small device helpers shared by all kernels.
// B200 (sm_100a) only; there is deliberately no CPU path in this library.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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
  cudaStream_t scanStream = nullptr;   // K1 launches of a streaming scan (overlaps the uploads queued on `stream`)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evMid = nullptr;   // evMid: end of the K1 launches of a punch-in scan
  int64_t launches = 0;        // total kernels launched on this context
  int64_t callLaunches0 = 0;   // snapshot at begin_call
  float lastMs = 0.f;
  int64_t lastLaunches = 0;
  int lastSelfKernel = 0;      // which kernel rendered the last SelfSimilarity image (sgz_self_last_kernel)

  int bind() const {
    SGZ_CUDA(cudaSetDevice(device));
    return SGZ_OK;
  }
  int begin_call() {
    callLaunches0 = launches;
    SGZ_CUDA(cudaEventRecord(ev0, stream));
    return SGZ_OK;
  }
  // the same without the host wait: the caller synchronises the stream anyway (a download follows) and then collects
  int end_call_async() {
    SGZ_CUDA(cudaEventRecord(ev1, stream));
    return SGZ_OK;
  }
  int collect_call() {
    SGZ_CUDA(cudaEventElapsedTime(&lastMs, ev0, ev1));
    lastLaunches = launches - callLaunches0;
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

// ---------------------------------------------------------------------------------------------
// Device memory pool.  cudaMalloc / cudaFree of the multi-GB database and curve buffers cost 10-100 ms
// each and serialise the whole device (measured with tools/e2e_probe.py), which a search that is
// re-created per query pays every time.  Freed blocks are therefore parked per device and handed out
// again to requests of a similar size; sgz_ctx_trim (or memory pressure) returns them to the driver.
// Blocks enter the pool only after a device-wide synchronisation, the guarantee cudaFree gave before.
// ---------------------------------------------------------------------------------------------
#include <map>
#include <mutex>

namespace sgz {

class DevicePool {
 public:
  static DevicePool &of(int device) {
    static std::mutex m;
    static std::map<int, DevicePool *> pools;
    std::lock_guard<std::mutex> g(m);
    DevicePool *&p = pools[device];
    if (!p) p = new DevicePool();   // lives as long as the process, like the CUDA primary context
    return *p;
  }
  static size_t round_up(size_t bytes) {
    const size_t q = bytes >= (1u << 20) ? (size_t)2 << 20 : 512;
    return (bytes + q - 1) / q * q;
  }
  cudaError_t get(size_t bytes, void **out, size_t *got) {
    const size_t want = round_up(bytes);
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.lower_bound(want);
      if (it != free_.end() && it->first <= want + want / 4) {   // at most 25 % internal waste
        *out = it->second;
        *got = it->first;
        cached_ -= it->first;
        free_.erase(it);
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {
      (void)cudaGetLastError();
      trim();
      e = cudaMalloc(out, want);
    }
    *got = want;
    return e;
  }
  void put(void *p, size_t bytes) {
    cudaDeviceSynchronize();   // nothing in flight may still touch the block
    std::lock_guard<std::mutex> g(mu_);
    if (cached_ + bytes > limit_) {
      cudaFree(p);
      return;
    }
    free_.emplace(bytes, p);
    cached_ += bytes;
  }
  size_t trim() {
    std::lock_guard<std::mutex> g(mu_);
    size_t n = cached_;
    for (auto &kv : free_) cudaFree(kv.second);
    free_.clear();
    cached_ = 0;
    return n;
  }
  size_t cached() {
    std::lock_guard<std::mutex> g(mu_);
    return cached_;
  }

 private:
  DevicePool() {
    if (const char *e = getenv("SGZ_POOL_MAX_GB")) limit_ = (size_t)(atof(e) * (double)(1ull << 30));
  }
  std::mutex mu_;
  std::multimap<size_t, void *> free_;
  size_t cached_ = 0;
  size_t limit_ = (size_t)64 << 30;   // of 180 GB; the 1000 h database + its curves are 20 GB
};

}  // namespace sgz

// simple owning device buffer (pool backed)
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;          // elements requested
  size_t bytes = 0;      // size of the pool block
  int dev = -1;
  int alloc(size_t count) {
    if (count <= n && p) return SGZ_OK;
    release();
    if (count == 0) return SGZ_OK;
    cudaGetDevice(&dev);
    void *q = nullptr;
    cudaError_t e = ::sgz::DevicePool::of(dev).get(count * sizeof(T), &q, &bytes);
    if (e != cudaSuccess) {
      p = nullptr;
      n = 0;
      bytes = 0;
      ::sgz::set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(e));
      (void)cudaGetLastError();
      return e == cudaErrorMemoryAllocation ? SGZ_ERR_NOMEM : SGZ_ERR_CUDA;
    }
    p = (T *)q;
    n = count;
    // test hook: hand out NaN-patterned memory so that a kernel relying on zero-initialised buffers shows up
    static const bool poison = getenv("SGZ_POOL_POISON") != nullptr;
    if (poison) {
      cudaMemset(q, 0xFF, bytes);
      cudaDeviceSynchronize();   // the NULL-stream memset is not ordered with the non-blocking streams
    }
    return SGZ_OK;
  }
  void release() {
    if (p) ::sgz::DevicePool::of(dev).put(p, bytes);
    p = nullptr;
    n = 0;
    bytes = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};
