// stats.cuh -- K6: FeatureStats (FeatureStatsImpl.scala:30-135), SURVEY.md section 8(f) rank 2: the producer of
// feat_norms.aif.  Per file and channel: min / max / mean of the RAW features, a 2048-bin histogram of the
// skew-warped values, the 1st and 99th percentile read off it; per channel over the database: the smallest 1st
// and the largest 99th percentile.
//
// Works on a database created with norm = NULL (raw values are stored bit-exactly), so the upload path -- big
// endian AIFF payloads, staging ring, device-side transposition -- is shared with the search.
//   pass 1  k_stats_minmax : one warp per (file, channel pair); coalesced loads, per-lane min / max, the Double sum
//                            in the reference's frame order (:70-84) by one lane per channel; ends with the skew (:86-92)
//   pass 2  k_stats_hist   : one block per (file, pair, 8192-frame segment); bins in shared memory, integer
//                            counts merged with atomics (order independent, exact) (:94-113)
//   pass 3  k_stats_pctl   : one warp per (file, channel) finds where the bin counts reach 1 % / 99 % (:115-131)
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

// One warp per (file, channel pair).  The Double sum must keep the reference's frame order (:70-84), so it is a
// chain of n dependent additions whatever the layout; everything around it is made parallel: the warp fetches 256
// frames with coalesced loads (the next 256 are in flight while the chain runs), min / max are taken per lane and
// merged at the end ("first one wins" among equal values, like the sequential scan, so that the sign of a zero
// extreme is the reference's), and lanes 0 / 1 add the two channels of the pair from shared memory.
constexpr int kMmWarps = 4;
constexpr int kMmSteps = 8;                       // 32-frame steps per chunk
constexpr int kMmChunk = 32 * kMmSteps;

struct StatExt {          // running extreme of one channel: value and the frame it was first seen at
  float v;
  int64_t at;
};

__device__ __forceinline__ void stat_ext_merge(StatExt &a, bool isMin, unsigned full, int delta) {
  const float ov = __shfl_xor_sync(full, a.v, delta);
  const int64_t oat = __shfl_xor_sync(full, a.at, delta);
  const bool better = isMin ? ov < a.v : ov > a.v;
  if (better || (ov == a.v && oat < a.at)) { a.v = ov; a.at = oat; }
}

__global__ void __launch_bounds__(32 * kMmWarps) k_stats_minmax(const StatsParams p) {
  __shared__ double buf[kMmWarps][2][kMmChunk];   // the chunk as Doubles, one plane per channel of the pair
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int idx = blockIdx.x * kMmWarps + warp;
  if (idx >= p.numFiles * p.numPairs) return;
  const int f = idx / p.numPairs, pr = idx - f * p.numPairs;
  const int64_t g0 = p.fileStart[f], n = p.fileStart[f + 1] - g0;
  const float2 *row = p.data + (int64_t)pr * p.rowStride + g0;
  const int64_t never = INT64_MAX;
  StatExt mn0{INFINITY, never}, mn1{INFINITY, never}, mx0{-INFINITY, never}, mx1{-INFINITY, never};
  double s = 0.0;                                  // lane 0: channel 2 pr, lane 1: channel 2 pr + 1
  float2 v[kMmSteps];
#pragma unroll
  for (int u = 0; u < kMmSteps; u++) {
    const int64_t i = 32 * u + lane;
    v[u] = i < n ? __ldg(row + i) : make_float2(0.f, 0.f);
  }
  const double *bsum = buf[warp][lane & 1];
  for (int64_t base = 0; base < n; base += kMmChunk) {
#pragma unroll
    for (int u = 0; u < kMmSteps; u++) {
      const int64_t i = base + 32 * u + lane;
      // Float -> Double by all lanes: the conversion unit is narrow, two lanes converting 2 x 256 values would
      // cost more than the additions themselves
      buf[warp][0][32 * u + lane] = (double)v[u].x;
      buf[warp][1][32 * u + lane] = (double)v[u].y;
      if (i < n) {
        if (v[u].x < mn0.v) { mn0.v = v[u].x; mn0.at = i; }
        if (v[u].x > mx0.v) { mx0.v = v[u].x; mx0.at = i; }
        if (v[u].y < mn1.v) { mn1.v = v[u].y; mn1.at = i; }
        if (v[u].y > mx1.v) { mx1.v = v[u].y; mx1.at = i; }
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < kMmSteps; u++) {           // next chunk: in flight during the chain
      const int64_t i = base + kMmChunk + 32 * u + lane;
      v[u] = i < n ? __ldg(row + i) : make_float2(0.f, 0.f);
    }
    const int cnt = (int)(n - base < kMmChunk ? n - base : kMmChunk);
    if (lane < 2) {
      int k = 0;
      for (; k + 8 <= cnt; k += 8) {
        double d[8];
#pragma unroll
        for (int q = 0; q < 8; q++) d[q] = bsum[k + q];
#pragma unroll
        for (int q = 0; q < 8; q++) s = __dadd_rn(s, d[q]);
      }
      for (; k < cnt; k++) s = __dadd_rn(s, bsum[k]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int delta = 16; delta >= 1; delta >>= 1) {
    stat_ext_merge(mn0, true, full, delta);
    stat_ext_merge(mn1, true, full, delta);
    stat_ext_merge(mx0, false, full, delta);
    stat_ext_merge(mx1, false, full, delta);
  }
  if (lane < 2) {
    const int c = 2 * pr + lane;
    if (c < p.numCh) {
      const float mn = lane ? mn1.v : mn0.v, mx = lane ? mx1.v : mx0.v;
      const double mean = __ddiv_rn(s, (double)n);
      const float d = __fsub_rn(mx, mn);
      const double m = __ddiv_rn(__dsub_rn(mean, (double)mn), (double)d);
      p.mins[(int64_t)f * p.numCh + c] = mn;
      p.maxs[(int64_t)f * p.numCh + c] = mx;
      p.skews[(int64_t)f * p.numCh + c] = __ddiv_rn(log(0.5), log(m));
    }
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

// One warp per (file, channel).  The reference walks the bins until the running count reaches 1 % of the frames and
// goes on from there to 99 % (:115-131): i = smallest index whose prefix count (bins 0 .. i-1) reaches the target, the
// second index never below the first.  Lane l owns bins [64 l, 64 l + 64): a warp scan of the lane sums finds the lane
// in which the target is crossed, that lane walks its 64 bins.
__device__ __forceinline__ int stat_bins_until(const int32_t *cp, int target, int lane, int laneSum, int laneIncl) {
  const unsigned full = 0xffffffffu;
  if (target <= 0) return 0;
  const unsigned reached = __ballot_sync(full, laneIncl >= target);
  if (reached == 0u) return kStatBins;
  const int l = __ffs(reached) - 1;
  int idx = 0;
  if (lane == l) {
    int acc = laneIncl - laneSum;
    for (int j = 0; j < kStatBins / 32; j++) {
      acc += cp[l * (kStatBins / 32) + j];
      if (acc >= target) { idx = l * (kStatBins / 32) + j + 1; break; }
    }
  }
  return __shfl_sync(full, idx, l);
}

__global__ void k_stats_pctl(const StatsParams p) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (idx >= p.fileCount * p.numCh) return;
  const int fl = idx / p.numCh, c = idx - fl * p.numCh, f = p.file0 + fl;
  const int64_t n = p.fileStart[f + 1] - p.fileStart[f];
  const int32_t *cp = p.hist + ((int64_t)fl * p.numCh + c) * kStatBins;
  const int p01n = j_d2i(__dmul_rn((double)n, 0.01));
  const int p99n = j_d2i(__dmul_rn((double)n, 0.99));
  int laneSum = 0;
  for (int j = 0; j < kStatBins / 32; j++) laneSum += cp[lane * (kStatBins / 32) + j];
  int laneIncl = laneSum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(full, laneIncl, d);
    if (lane >= d) laneIncl += o;
  }
  const int i01 = stat_bins_until(cp, p01n, lane, laneSum, laneIncl);
  const int i99 = max(i01, stat_bins_until(cp, p99n, lane, laneSum, laneIncl));
  if (lane == 0) {
    const double skewr = __ddiv_rn(1.0, p.skews[(int64_t)f * p.numCh + c]);
    const float mn = p.mins[(int64_t)f * p.numCh + c];
    const float d = __fsub_rn(p.maxs[(int64_t)f * p.numCh + c], mn);
    p.perFile[((int64_t)f * p.numCh + c) * 2] =
        __dadd_rn(__dmul_rn(pow(__ddiv_rn((double)i01, 2048.0), skewr), (double)d), (double)mn);
    p.perFile[((int64_t)f * p.numCh + c) * 2 + 1] =
        __dadd_rn(__dmul_rn(pow(__ddiv_rn((double)i99, 2048.0), skewr), (double)d), (double)mn);
  }
}

}  // namespace sgz
