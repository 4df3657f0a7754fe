// peaks.cuh -- live micro-benchmarks of the pipes our kernels are bound by.  K1 (sliding
// correlation at W = 172) is FP32-FFMA bound, and MEASURED_PEAKS.json only carries HBM and bf16
// tensor figures, so bench.py measures the FFMA roofline on the same GPU, in the same process,
// right before it times K1 ("FP32 peak must be measured").
#pragma once
#include "common.cuh"

namespace sgz {

template <int ILP>
__global__ void k_peak_ffma(float *out, int iters, float b, float c) {
  float a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void k_peak_ffma2(float *out, int iters, float b, float c) {
  unsigned long long a[ILP], bb, cc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    float x = (float)(threadIdx.x + i) * 1e-3f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(x), "f"(x + 1.f));
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb), "l"(cc));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; i++) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi;
  }
  if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void k_peak_dfma(double *out, int iters, double b, double c) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  if (s == 12345.678) out[0] = s;
}

// Register-operand bound patterns (what a register-tiled FP32 kernel really issues): an 8x8 outer
// product c[i][j] += a[i] * b[j] where a is reused across a row (reuse cache) but b[j] and c[i][j] are
// fresh register-file reads for every FFMA.
__global__ void k_peak_ffma_outer(float *out, int iters, float seed) {
  float a[8], b[8], c[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = out[8 + i] + seed;      // runtime values: nothing folds into immediates
    b[i] = out[16 + i] - seed;
#pragma unroll
    for (int j = 0; j < 8; j++) c[i][j] = 0.f;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) c[i][j] = fmaf(a[i], b[j], c[i][j]);
    }
    // keep a/b live and changing without adding arithmetic to the measured pattern's ratio
    a[it & 7] += 1e-6f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) s += c[i][j];
  }
  if (s == 12345.678f) out[0] = s;
}

// same outer product with packed FFMA2: c2[i][j] (pair along j) += (a[i], a[i]) * (b[2j], b[2j+1])
__global__ void k_peak_ffma2_outer(float *out, int iters, float seed) {
  unsigned long long a2[8], b2[4], c2[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float x = out[8 + i] + seed;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a2[i]) : "f"(x), "f"(x));
#pragma unroll
    for (int j = 0; j < 4; j++) asm("mov.b64 %0, {%1, %2};" : "=l"(c2[i][j]) : "f"(0.f), "f"(0.f));
  }
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float y = out[16 + 2 * j] - seed, y2 = out[17 + 2 * j] - seed;
    asm("mov.b64 %0, {%1, %2};" : "=l"(b2[j]) : "f"(y), "f"(y2));
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c2[i][j]) : "l"(a2[i]), "l"(b2[j]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(c2[i][j]));
      s += lo + hi;
    }
  }
  if (s == 12345.678f) out[0] = s;
}

__global__ void k_peak_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

__global__ void k_peak_lds(float *out, int iters) {
  __shared__ float4 sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float4 v = sh[(idx + 32 * k) & 1023];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx += 3;
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

inline int measure_peak(sgz_ctx *ctx, int which, double *value) {
  SGZ_TRY(ctx->bind());
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;
  DevBuf<float> sink;
  SGZ_TRY(sink.alloc(64));
  const int blocks = ctx->smCount * 8, threads = 256;
  float ms = 0.f;
  double best = 0.0;
  if (which >= 0 && which <= 2) {
    const int iters = 4096;
    constexpr int ILP = 16;
    for (int rep = 0; rep < 4; rep++) {
      SGZ_CUDA(cudaEventRecord(e0, st));
      if (which == 0) k_peak_ffma<ILP><<<blocks, threads, 0, st>>>(sink.p, iters, 0.999f, 1e-3f);
      else if (which == 1) k_peak_ffma2<ILP><<<blocks, threads, 0, st>>>(sink.p, iters, 0.999f, 1e-3f);
      else k_peak_dfma<ILP><<<blocks, threads, 0, st>>>((double *)sink.p, iters, 0.999, 1e-3);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_CUDA(cudaEventRecord(e1, st));
      SGZ_CUDA(cudaEventSynchronize(e1));
      SGZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      double flops = 2.0 * (double)blocks * threads * iters * ILP * (which == 1 ? 2.0 : 1.0);
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
  } else if (which == 5 || which == 6) {
    const int iters = 2048;
    for (int rep = 0; rep < 4; rep++) {
      SGZ_CUDA(cudaEventRecord(e0, st));
      if (which == 5) k_peak_ffma_outer<<<blocks, threads, 0, st>>>(sink.p, iters, 0.5f);
      else k_peak_ffma2_outer<<<blocks, threads, 0, st>>>(sink.p, iters, 0.5f);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_CUDA(cudaEventRecord(e1, st));
      SGZ_CUDA(cudaEventSynchronize(e1));
      SGZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      double flops = 2.0 * (double)blocks * threads * iters * 64.0;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
  } else if (which >= 10 && which < 20) {
    // FFMA2 outer product with (which - 9) warps per scheduler (one CTA per SM): issue-interval probe
    const int warpsPerSmsp = which - 9;
    const int iters = 8192;
    for (int rep = 0; rep < 4; rep++) {
      SGZ_CUDA(cudaEventRecord(e0, st));
      k_peak_ffma2_outer<<<ctx->smCount, 128 * warpsPerSmsp, 0, st>>>(sink.p, iters, 0.5f);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_CUDA(cudaEventRecord(e1, st));
      SGZ_CUDA(cudaEventSynchronize(e1));
      SGZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      double flops = 2.0 * (double)ctx->smCount * 128 * warpsPerSmsp * iters * 64.0;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
  } else if (which == 3) {
    const size_t n = (size_t)1 << 27;  // 2 GiB per buffer as float4
    DevBuf<float4> a, b;
    SGZ_TRY(a.alloc(n));
    SGZ_TRY(b.alloc(n));
    SGZ_CUDA(cudaMemsetAsync(a.p, 0, n * sizeof(float4), st));
    for (int rep = 0; rep < 4; rep++) {
      SGZ_CUDA(cudaEventRecord(e0, st));
      k_peak_copy<<<ctx->smCount * 16, 512, 0, st>>>(a.p, b.p, n);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_CUDA(cudaEventRecord(e1, st));
      SGZ_CUDA(cudaEventSynchronize(e1));
      SGZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0) best = std::max(best, 2.0 * n * sizeof(float4) / (ms * 1e-3) / 1e9);
    }
  } else if (which == 4) {
    const int iters = 4096;
    for (int rep = 0; rep < 4; rep++) {
      SGZ_CUDA(cudaEventRecord(e0, st));
      k_peak_lds<<<blocks, threads, 0, st>>>(sink.p, iters);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_CUDA(cudaEventRecord(e1, st));
      SGZ_CUDA(cudaEventSynchronize(e1));
      SGZ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0) best = std::max(best, (double)blocks * threads * iters * 8 * 16.0 / (ms * 1e-3) / 1e9);
    }
  } else {
    set_error("sgz_measure_peak: unknown pipe %d", which);
    return SGZ_ERR_INVALID;
  }
  *value = best;
  return SGZ_OK;
}

}  // namespace sgz
