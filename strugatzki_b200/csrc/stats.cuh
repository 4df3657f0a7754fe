// stats.cuh -- K6: FeatureStats (FeatureStatsImpl.scala:30-135), SURVEY.md section 8(f) rank 2: the producer of
// feat_norms.aif.  Per file and channel: min / max / mean of the RAW features, a 2048-bin histogram of the
// skew-warped values, the 1st and 99th percentile read off it; per channel over the database: the smallest 1st
// and the largest 99th percentile.
//
// Works on a database created with norm = NULL (raw values are stored bit-exactly), so the upload path -- big
// endian AIFF payloads, staging ring, device-side transposition -- is shared with the search.
//   pass 1  k_stats_minmax : one thread per (file, channel pair) walks the file in frame order, so the Double sum
//                            has the reference's summation order (:70-84); ends with the skew (:86-92)
//   pass 2  k_stats_hist   : one block per (file, pair, 8192-frame segment); bins in shared memory, integer
//                            counts merged with atomics (order independent, exact) (:94-113)
//   pass 3  k_stats_pctl   : one thread per (file, channel) scans its 2048 bins (:115-131)
// HBM-bound by design (2 x 56 B per frame); pass 2 carries one FP64 pow per value.
#pragma once
#include "common.cuh"

namespace sgz {

constexpr int kStatBins = 2048;
constexpr int kStatSeg = 8192;   // frames per histogram block

struct StatsParams {
  const float2 *data;        // pair rows [numPairs][rowStride], RAW values
  int64_t rowStride;
  const int64_t *fileStart;  // [numFiles+1]
  int numFiles, numCh, numPairs;
  int file0, fileCount;      // batch of files handled by this launch (histogram memory is per batch)
  float *mins, *maxs;        // [numFiles][numCh]
  double *skews;             // [numFiles][numCh]
  int32_t *hist;             // [fileCount][numCh][kStatBins]
  double *perFile;           // [numFiles][numCh][2] = (p01, p99)
};

// Double -> Int like the JVM: NaN -> 0, saturating, truncating
__device__ __forceinline__ int j_d2i(double x) { return x != x ? 0 : __double2int_rz(x); }

__global__ void k_stats_minmax(const StatsParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.numFiles * p.numPairs) return;
  const int f = idx / p.numPairs, pr = idx - f * p.numPairs;
  const int64_t g0 = p.fileStart[f], n = p.fileStart[f + 1] - g0;
  const float2 *row = p.data + (int64_t)pr * p.rowStride + g0;
  float mn0 = INFINITY, mn1 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY;
  double s0 = 0.0, s1 = 0.0;
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {     // 8 independent loads in flight, the adds stay in frame order
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = __ldg(row + i + k);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (v[k].x < mn0) mn0 = v[k].x;
      if (v[k].x > mx0) mx0 = v[k].x;
      s0 = __dadd_rn(s0, (double)v[k].x);
      if (v[k].y < mn1) mn1 = v[k].y;
      if (v[k].y > mx1) mx1 = v[k].y;
      s1 = __dadd_rn(s1, (double)v[k].y);
    }
  }
  for (; i < n; i++) {
    const float2 v = __ldg(row + i);
    if (v.x < mn0) mn0 = v.x;
    if (v.x > mx0) mx0 = v.x;
    s0 = __dadd_rn(s0, (double)v.x);
    if (v.y < mn1) mn1 = v.y;
    if (v.y > mx1) mx1 = v.y;
    s1 = __dadd_rn(s1, (double)v.y);
  }
  const double log05 = log(0.5);
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int c = 2 * pr + h;
    if (c >= p.numCh) break;
    const float mn = h ? mn1 : mn0, mx = h ? mx1 : mx0;
    const double mean = __ddiv_rn(h ? s1 : s0, (double)n);
    const float d = __fsub_rn(mx, mn);
    const double m = __ddiv_rn(__dsub_rn(mean, (double)mn), (double)d);
    p.mins[(int64_t)f * p.numCh + c] = mn;
    p.maxs[(int64_t)f * p.numCh + c] = mx;
    p.skews[(int64_t)f * p.numCh + c] = __ddiv_rn(log05, log(m));
  }
}

// grid = (segments, numPairs, fileCount)
__global__ void k_stats_hist(const StatsParams p) {
  __shared__ int32_t sh[2 * kStatBins];
  const int f = p.file0 + blockIdx.z, pr = blockIdx.y;
  const int64_t g0 = p.fileStart[f], n = p.fileStart[f + 1] - g0;
  const int64_t s0 = (int64_t)blockIdx.x * kStatSeg;
  if (s0 >= n) return;
  const int64_t s1 = min(n, s0 + kStatSeg);
  for (int i = threadIdx.x; i < 2 * kStatBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int c0 = 2 * pr, c1 = 2 * pr + 1;
  const bool has1 = c1 < p.numCh;
  const float mn0 = p.mins[(int64_t)f * p.numCh + c0], d0 = __fsub_rn(p.maxs[(int64_t)f * p.numCh + c0], mn0);
  const double k0 = p.skews[(int64_t)f * p.numCh + c0];
  const float mn1 = has1 ? p.mins[(int64_t)f * p.numCh + c1] : 0.f;
  const float d1 = has1 ? __fsub_rn(p.maxs[(int64_t)f * p.numCh + c1], mn1) : 1.f;
  const double k1 = has1 ? p.skews[(int64_t)f * p.numCh + c1] : 1.0;
  const float2 *row = p.data + (int64_t)pr * p.rowStride + g0;
  for (int64_t i = s0 + threadIdx.x; i < s1; i += blockDim.x) {
    const float2 v = __ldg(row + i);
    // (math.pow((f - min) / d, skew) * 2047 + 0.5).toInt -- Float sub / div, Double pow, no FMA, truncation
    int b = j_d2i(__dadd_rn(__dmul_rn(pow((double)__fdiv_rn(__fsub_rn(v.x, mn0), d0), k0), 2047.0), 0.5));
    if (b >= 0 && b < kStatBins) atomicAdd(&sh[b], 1);
    if (has1) {
      b = j_d2i(__dadd_rn(__dmul_rn(pow((double)__fdiv_rn(__fsub_rn(v.y, mn1), d1), k1), 2047.0), 0.5));
      if (b >= 0 && b < kStatBins) atomicAdd(&sh[kStatBins + b], 1);
    }
  }
  __syncthreads();
  int32_t *h0 = p.hist + ((int64_t)blockIdx.z * p.numCh + c0) * kStatBins;
  for (int i = threadIdx.x; i < (has1 ? 2 : 1) * kStatBins; i += blockDim.x)   // channels c0, c1 are adjacent
    if (sh[i]) atomicAdd(h0 + i, sh[i]);
}

__global__ void k_stats_pctl(const StatsParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.fileCount * p.numCh) return;
  const int fl = idx / p.numCh, c = idx - fl * p.numCh, f = p.file0 + fl;
  const int64_t n = p.fileStart[f + 1] - p.fileStart[f];
  const int32_t *cp = p.hist + ((int64_t)fl * p.numCh + c) * kStatBins;
  const int p01n = j_d2i(__dmul_rn((double)n, 0.01));
  const int p99n = j_d2i(__dmul_rn((double)n, 0.99));
  const double skewr = __ddiv_rn(1.0, p.skews[(int64_t)f * p.numCh + c]);
  const float mn = p.mins[(int64_t)f * p.numCh + c];
  const float d = __fsub_rn(p.maxs[(int64_t)f * p.numCh + c], mn);
  int cnt = 0, i = 0;
  while (cnt < p01n && i < kStatBins) { cnt += cp[i]; i++; }
  p.perFile[((int64_t)f * p.numCh + c) * 2] =
      __dadd_rn(__dmul_rn(pow(__ddiv_rn((double)i, 2048.0), skewr), (double)d), (double)mn);
  while (cnt < p99n && i < kStatBins) { cnt += cp[i]; i++; }
  p.perFile[((int64_t)f * p.numCh + c) * 2 + 1] =
      __dadd_rn(__dmul_rn(pow(__ddiv_rn((double)i, 2048.0), skewr), (double)d), (double)mn);
}

}  // namespace sgz
