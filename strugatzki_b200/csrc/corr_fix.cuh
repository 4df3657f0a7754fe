// corr_fix.cuh -- exact re-evaluation of the ill-conditioned offsets of K1.
//
// The tensor-core K1 (corr_tc2.cuh) gets the window variance as E[x^2] - E[x]^2 from FP64 sums and the cross term from
// split-FP16 products whose accumulation truncates at the scale of the LEVEL of the data: both lose accuracy relative to
// the spread when a window is almost constant (variance below 2e-3 of the mean square, i.e. a spread below 4.5 % of the
// level -- digital silence, a held tone; measured: 2.8e-6 absolute at 1e-3).  The reference has no such limit:
// MathUtil.stat / correlate run two-pass in Double and return NaN only for an EXACTLY constant window (0 / 0,
// MathUtil.scala:29-62,177-196).  K1 therefore only flags such offsets (a sentinel NaN in the curve and an entry in a
// list); this kernel replays the reference's Double arithmetic for them operation by operation -- stat in the physical
// order of the ring buffer (frame j of the file sits at j % W, FeatureCorrelationImpl.scala:190-246), correlate in
// logical order, no FMA contraction, like K3 / K5 -- so their sims are the oracle's bit for bit: NaN exactly where the
// reference says NaN, finite and accurate for near-constant windows.
#pragma once
#include "common.cuh"

namespace sgz {

constexpr uint32_t kFixSentinel = 0x7fc00badu;   // quiet NaN with a payload: "ill-conditioned, to be re-evaluated"

// calcBoost (FeatureCorrelationImpl.scala:73-78) of the window that starts at global frame g (file-local offset tl).
// The tensor-core K1 only evaluates the gate `boost <= maxBoost` per offset (a comparison of the loudness sum); the boost
// VALUE is needed for the handful of offsets that end up in a result or are asked for, and is computed here like the
// reference does: MathUtil.avg over ring positions 0 .. W-1 in Double -> Float, then exp((ln avgIn - ln avg) / 0.6).
// The round-1 kernels still write a boost curve (`arr`).
struct BoostSrc {
  const float *arr;         // boost curve of the scan, or nullptr: compute from the loudness channel
  const float2 *data;       // pair row 0 holds (loudness, first spectral channel)
  int W;
  double lnAvgIn;
  __device__ float at(int64_t g, int64_t tl) const {
    if (arr) return arr[g];
    const int r0 = (int)(tl % W);
    double sum = 0.0;
    for (int pos = 0; pos < W; pos++) {
      const int i = pos >= r0 ? pos - r0 : pos - r0 + W;
      sum = __dadd_rn(sum, (double)data[g + i].x);
    }
    const float avg = (float)__ddiv_rn(sum, (double)W);
    return (float)exp(__ddiv_rn(__dsub_rn(lnAvgIn, log((double)avg)), 0.6));
  }
};

struct CorrFixParams {
  const float2 *data;       // normalised pair rows [numPairs][rowStride]
  int64_t rowStride, usedFrames;
  const double *a;          // centred query [numCh][W]: (double)a[c][i] + (-mean of its group), first factor of correlate
  int numCh, W;
  double stdT, stdS;        // MathUtil.stat of the query groups
  float weight, maxBoost;
  const int64_t *fileStart;
  int numFiles, tailExtra;
  const uint32_t *list;     // flagged global offsets (window start frames)
  const uint32_t *count;    // number of flagged offsets; > cap = the list overflowed: scan the curve for sentinels
  uint32_t cap;
  float *sim;
  BoostSrc boost;
  unsigned long long *fileMax;
  uint32_t *fileNaN;        // [numFiles] set when a window of the file is NaN in the reference (exactly constant), or nullptr
  int listOnly;             // 1: a list that overflowed is left alone (corr_refine.cuh) instead of scanning the curve for sentinels
};

__device__ __forceinline__ float fix_value(const CorrFixParams &p, int ch, int64_t g) {
  const float2 v = p.data[(int64_t)(ch >> 1) * p.rowStride + g];
  return (ch & 1) ? v.y : v.x;
}

// FeatureCorrelationImpl.correlate (:414-421) of the query group [chanOff, chanOff + numChannels) against the window
// that starts at global frame g (file-local offset tl): stat over ring positions 0 .. W-1, then MathUtil.correlate
__device__ float fix_correlate(const CorrFixParams &p, int64_t g, int64_t tl, int chanOff, int numChannels, double aStd) {
  const int W = p.W;
  const int matSize = numChannels * W;
  const int r0 = (int)(tl % W);              // ring position of the window's first frame
  // ring position pos holds window frame (pos - r0) mod W
  double sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    for (int pos = 0; pos < W; pos++) {
      const int i = pos >= r0 ? pos - r0 : pos - r0 + W;
      sum = __dadd_rn(sum, (double)fix_value(p, ch + chanOff, g + i));
    }
  }
  const double bMean = __ddiv_rn(sum, (double)matSize);
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    for (int pos = 0; pos < W; pos++) {
      const int i = pos >= r0 ? pos - r0 : pos - r0 + W;
      const double d = __dsub_rn((double)fix_value(p, ch + chanOff, g + i), bMean);
      sum = __dadd_rn(sum, __dmul_rn(d, d));
    }
  }
  const double bStd = __dsqrt_rn(__ddiv_rn(sum, (double)matSize));
  const double bAdd = -bMean;
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const double *ca = p.a + (int64_t)(ch + chanOff) * W;
    for (int i = 0; i < W; i++)
      sum = __dadd_rn(sum, __dmul_rn(ca[i], __dadd_rn((double)fix_value(p, ch + chanOff, g + i), bAdd)));
  }
  return (float)__ddiv_rn(sum, __dmul_rn(__dmul_rn(aStd, bStd), (double)matSize));
}

__device__ void fix_one(const CorrFixParams &p, int64_t g) {
  int lo = 0, hi = p.numFiles;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.fileStart[mid] <= g) lo = mid; else hi = mid;
  }
  const int64_t tl = g - p.fileStart[lo];
  const float boost = p.boost.at(g, tl);
  float sim = 0.f;
  if (boost <= p.maxBoost) {                 // :199 (NaN boost: the comparison is false like the reference's)
    const float temporal = p.weight > 0.f ? fix_correlate(p, g, tl, 0, 1, p.stdT) : 0.f;
    const float spectral = p.weight < 1.f ? fix_correlate(p, g, tl, 1, p.numCh - 1, p.stdS) : 0.f;
    sim = __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
  }
  p.sim[g] = sim;
  if (sim != sim && p.fileNaN) p.fileNaN[lo] = 1u;
  if (sim == sim && p.fileMax) {
    const unsigned long long key = ((unsigned long long)float_order_key(sim) << 32) |
                                   (unsigned long long)(0xffffffffu - (uint32_t)tl);
    atomicMax(p.fileMax + lo, key);
  }
}

// boost of the offset that holds each file's maximum (fileMax key: order key of the sim << 32 | ~offset; 0 = no offset).
// With numPerFile = 1 the entry of a file IS that maximum, so a punch-in search needs no selection kernel at all.
__global__ void k_filemax_boost(BoostSrc b, const int64_t *__restrict__ fileStart, const unsigned long long *__restrict__ fileMax,
                                int numFiles, float *__restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= numFiles) return;
  const unsigned long long key = fileMax[f];
  if (key == 0ull) { out[f] = 1.0f; return; }
  const int64_t tl = (int64_t)(0xffffffffu - (uint32_t)key);
  out[f] = b.at(fileStart[f] + tl, tl);
}

// Boost CURVE of a punch window for every global frame (punch-out searches: the grid kernels look a boost up per candidate
// CELL, thousands per file, where a punch-in search needs a handful).  A thread slides the Double loudness sum over a run
// of 32 offsets; the sum differs from the reference's ring-order sum in its last bits only (the boost is compared at 1e-5).
constexpr int kBoostRun = 32;
__global__ void k_boost_all(const float2 *__restrict__ data, int64_t usedFrames, int W, double lnAvgIn, float *__restrict__ out) {
  const int64_t g0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * kBoostRun;
  if (g0 >= usedFrames) return;
  double sum = 0.0;
  for (int i = 0; i < W; i++) sum += (double)data[g0 + i].x;          // (the zero slack behind the database is readable)
  const int n = (int)min((int64_t)kBoostRun, usedFrames - g0);
  for (int j = 0; j < n; j++) {
    const float avg = (float)(sum / (double)W);
    out[g0 + j] = expf(((float)lnAvgIn - logf(avg)) * (1.0f / 0.6f));      // (1e-6 relative; the contract is 1e-5)
    sum += (double)data[g0 + j + W].x - (double)data[g0 + j].x;
  }
}

// boost values of n consecutive offsets of one file (sgz_corr_curve)
__global__ void k_boost_curve(BoostSrc b, int64_t g0, int64_t tl0, int64_t n, float *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = b.at(g0 + i, tl0 + i);
}

__global__ void __launch_bounds__(128) k_corr_fixup(const CorrFixParams p) {
  const uint32_t n = *p.count;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n <= p.cap) {
    for (int64_t i = t; i < (int64_t)n; i += stride) fix_one(p, (int64_t)p.list[i]);
  } else if (!p.listOnly) {
    for (int64_t g = t; g < p.usedFrames; g += stride)
      if (__float_as_uint(p.sim[g]) == kFixSentinel) fix_one(p, g);
  }
}

}  // namespace sgz
