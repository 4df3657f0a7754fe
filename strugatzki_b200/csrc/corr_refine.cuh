// corr_refine.cuh -- exact sims for the DECISIVE offsets of a punch-in search.
//
// The tensor-core scan (corr_tc2.cuh) is accurate to about 1e-6, the reference's queues (addMatch,
// FeatureCorrelationImpl.scala:120-150) compare exact Float sims: equal sims of repeated material are one entry of the
// TreeSet, a sim that differs in its last bit is a second one.  allPrio ends up with the numMatches largest entries, and
// the maximum of a file always survives in its entryPrio, so the final lowest sim is at least the numMatches-th largest
// DISTINCT file maximum T -- known after the scan.  Every offset whose sim reaches T - margin (margin = several times the
// kernel's error) is therefore re-evaluated with the reference's own Double arithmetic (fix_one, corr_fix.cuh: bit
// identical to the oracle), the curve is patched and the file maxima of the files concerned are rebuilt from the exact
// values (first position among exact ties) BEFORE any selection looks at them.  What a search returns is then what the
// reference returns, bit for bit, as long as the decisive offsets fit the list (kRefineCap; beyond it the approximate
// values stay, within the 1e-5 contract).  Fewer files with offsets than numMatches: T = -inf, everything is re-evaluated
// (small databases).  Punch-out searches (sims of CELLS, products of two curves) keep the approximate curves.
#pragma once
#include "corr_fix.cuh"

namespace sgz {

constexpr uint32_t kRefineCap = 1u << 18;
constexpr float kRefineMargin = 5e-5f;      // >= 2 x (1e-5 relative at |sim| <= 1 + 2e-6), the bound the parity tests hold K1 to
constexpr float kRefineTieTol = 2.5e-5f;    // file maxima closer than this may be equal in the reference

// thr[0] = a lower bound of the K-th largest DISTINCT file maximum, minus the margin; -inf when there are not that many.
// Distinct, because equal sims are one entry of the reference's TreeSet (a file that occurs twice does not take two places),
// and whether two approximate maxima are equal is only known after the re-evaluation: maxima closer than tieTol (twice the
// kernel's error) count as one.  K rounds of "largest key below the previous one minus tieTol" by one block; the keys are
// the high words of fileMax (order keys of the sims; 0 = file without an offset), held in registers for up to 8192 files
// (a round is then 32 compares, two warp reductions and ONE barrier).
// Then the files whose maximum reaches the threshold are listed (cand, candCount).
constexpr int kRefineMaxK = 4096;
constexpr int kRefineThrThreads = 256, kRefineThrRegs = 32;
__global__ void __launch_bounds__(kRefineThrThreads) k_refine_threshold(const unsigned long long *__restrict__ fileMax, int numFiles,
                                                                        int K, float margin, float tieTol, float *__restrict__ thr,
                                                                        int32_t *__restrict__ cand, uint32_t *__restrict__ candCount) {
  constexpr int kWarps = kRefineThrThreads / 32;
  __shared__ uint32_t warpMax[2][kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool inRegs = numFiles <= kRefineThrRegs * kRefineThrThreads;
  uint32_t keys[kRefineThrRegs];
#pragma unroll
  for (int j = 0; j < kRefineThrRegs; j++) {
    const int f = threadIdx.x + j * kRefineThrThreads;
    keys[j] = (inRegs && f < numFiles) ? (uint32_t)(fileMax[f] >> 32) : 0u;
  }
  uint32_t bound = 0xffffffffu, last = 0u;      // keys strictly below `bound` are still to be counted
  const int rounds = K > kRefineMaxK ? 0 : max(K, 1);
  for (int round = 0; round < rounds; round++) {
    uint32_t m = 0u;
    if (inRegs) {
#pragma unroll
      for (int j = 0; j < kRefineThrRegs; j++) {
        const uint32_t k = keys[j] < bound ? keys[j] : 0u;
        m = max(m, k);
      }
    } else {
      for (int f = threadIdx.x; f < numFiles; f += blockDim.x) {
        const uint32_t k = (uint32_t)(fileMax[f] >> 32);
        if (k < bound && k > m) m = k;
      }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) warpMax[round & 1][warp] = m;
    __syncthreads();                             // (the other buffer is rewritten in the next round, behind this barrier)
    m = __reduce_max_sync(0xffffffffu, lane < kWarps ? warpMax[round & 1][lane] : 0u);
    last = m;
    if (last == 0u) break;                       // fewer distinct maxima than K
    bound = float_order_key(float_from_order_key(last) - tieTol);
  }
  const float t = last == 0u ? -INFINITY : float_from_order_key(last) - margin;
  if (threadIdx.x == 0) thr[0] = t;
  for (int f = threadIdx.x; f < numFiles; f += blockDim.x) {
    const uint32_t k = (uint32_t)(fileMax[f] >> 32);
    if (k != 0u && float_from_order_key(k) >= t) cand[atomicAdd(candCount, 1u)] = f;
  }
}

// offsets of the listed files that reach the threshold.  Work item = (listed file, one of kRefineChunks pieces of its
// curve).  numPerFile = 1 (ownMax): the entry of a file is its maximum, so only offsets within the margin of the file's
// OWN maximum matter.
constexpr int kRefineChunks = 32;
constexpr size_t kExactBytesPerValue = 24;   // shared memory of k_corr_exact per window value and warp (rounded up from 20)
__global__ void __launch_bounds__(256) k_refine_collect(const float *__restrict__ sim, const int64_t *__restrict__ fileStart,
                                                        const unsigned long long *__restrict__ fileMax, int W, int tailExtra,
                                                        const float *__restrict__ thr, float margin, int ownMax,
                                                        const int32_t *__restrict__ cand, const uint32_t *__restrict__ candCount,
                                                        uint32_t *__restrict__ list, uint32_t *__restrict__ count, uint32_t cap) {
  const int64_t items = (int64_t)candCount[0] * kRefineChunks;
  float t0 = thr[0];
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int f = cand[item / kRefineChunks], piece = (int)(item % kRefineChunks);
    float t = t0;
    if (ownMax) t = fmaxf(t, float_from_order_key((uint32_t)(fileMax[f] >> 32)) - margin);
    const int64_t g0 = fileStart[f], n = (fileStart[f + 1] - g0) - tailExtra - W + 1;
    const int64_t per = (n + kRefineChunks - 1) / kRefineChunks, lo = piece * per, hi = min(n, lo + per);
    for (int64_t tl = lo + threadIdx.x; tl < hi; tl += 4 * blockDim.x) {
      float s[4];
#pragma unroll
      for (int j = 0; j < 4; j++) s[j] = tl + j * blockDim.x < hi ? sim[g0 + tl + j * blockDim.x] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        // (an exact 0 is the boost gate's answer, not an evaluation; NaN fails the comparison)
        if (s[j] >= t && s[j] != 0.f) {
          const uint32_t slot = atomicAdd(count, 1u);
          if (slot < cap) list[slot] = (uint32_t)(g0 + tl + j * blockDim.x);
        }
      }
    }
  }
}

// the exact maxima replace the scan's (files without a re-evaluated offset keep theirs)
__global__ void k_refine_merge(const unsigned long long *__restrict__ exact, unsigned long long *__restrict__ fileMax, int numFiles) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < numFiles && exact[f] != 0ull) fileMax[f] = exact[f];
}

// ---------------------------------------------------------------------------------------------
// fix_one (corr_fix.cuh) by a WARP: the lanes bring the window into shared memory with coalesced loads, then the serial
// Double chains of the reference -- which fix the order of every addition -- run side by side on a few lanes, from shared
// memory: phase 1 the ring-order sums of the temporal (= loudness: the boost's sum too) and the spectral group on two
// lanes, phase 2 variance and cross term of both groups on four.  The arithmetic is fix_correlate's operation for
// operation (a - b == a + (-b) in IEEE), so the result is the same bit pattern; the thread-per-offset kernel needed 0.45 ms
// for a handful of offsets (7 200 dependent global loads each).
// ---------------------------------------------------------------------------------------------
// sum = dadd(sum, f(i)) for i = lo .. hi-1, in that order; the operands of eight steps are fetched (shared memory,
// conversions) before the eight dependent additions, so the chain costs the latency of a DADD per element, not that of a
// load + conversion + DADD
template <class F>
__device__ __forceinline__ double ex_chain(double sum, int lo, int hi, F f) {
  int i = lo;
  for (; i + 8 <= hi; i += 8) {
    double d[8];
#pragma unroll
    for (int j = 0; j < 8; j++) d[j] = f(i + j);
#pragma unroll
    for (int j = 0; j < 8; j++) sum = __dadd_rn(sum, d[j]);
  }
  for (; i < hi; i++) sum = __dadd_rn(sum, f(i));
  return sum;
}

// ring position 0 holds window frame (W - r0) mod W: ring order = frames [W - r0, W) then [0, W - r0)
__device__ __forceinline__ double ex_ring_sum(const float *vals, int W, int r0, int chBegin, int nCh) {
  double sum = 0.0;
  const int split = r0 == 0 ? 0 : W - r0;
  for (int ch = chBegin; ch < chBegin + nCh; ch++) {
    const float *v = vals + ch * W;
    auto f = [&](int i) { return (double)v[i]; };
    sum = ex_chain(sum, split, W, f);
    sum = ex_chain(sum, 0, split, f);
  }
  return sum;
}

// The sum of n Floats in Double does not depend on the ORDER of the additions when no addition can round: every partial sum
// of any order is a multiple of the smallest ulp among the terms, and it is exactly representable when its magnitude stays
// below 2^53 of those ulps.  sum |v| bounds every partial sum, so with  sum |v| < 2^52 ulp_min  (one bit to spare for the
// rounding of the bound itself) the whole warp adds in parallel and gets the Double the reference's serial loop gets.
// Typical feature data (values of 1e-3 .. 1, a few thousand terms) needs 44 of the 52 bits.  Returns false when the
// condition fails (tiny or non-finite terms): the caller then runs the serial chain.
__device__ __forceinline__ bool ex_parallel_sum(const float *v, int n, int lane, double &sum) {
  double part = 0.0, mag = 0.0;
  uint32_t bmin = 255u;
  bool bad = false;
  for (int i = lane; i < n; i += 32) {
    const float x = v[i];
    const uint32_t b = (__float_as_uint(x) >> 23) & 255u;
    part += (double)x;
    mag += fabs((double)x);
    if (x != 0.f) { bmin = min(bmin, b); bad |= (b == 0u) | (b == 255u); }      // denormal / Inf / NaN: serial path
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, d);
    mag += __shfl_xor_sync(0xffffffffu, mag, d);
  }
  bmin = __reduce_min_sync(0xffffffffu, bmin);
  bad = __any_sync(0xffffffffu, bad);
  sum = part;
  if (bad) return false;
  if (bmin == 255u) return true;                                                // all terms zero
  return mag < ldexp(1.0, 52 + (int)bmin - 150);                                // ulp of a Float with exponent field b = 2^(b - 150)
}

// sum of n Doubles in shared memory, in order (eight loads, then the eight dependent additions)
__device__ __forceinline__ double ex_sum_doubles(const double *d, int n) {
  return ex_chain(0.0, 0, n, [&](int i) { return d[i]; });
}

__global__ void __launch_bounds__(128) k_corr_exact(const CorrFixParams p) {
  extern __shared__ double exTerms[];           // per warp: [2][numCh][W] Doubles (variance / cross terms), [numCh][W] Floats
  const uint32_t n = *p.count;
  if (n > p.cap) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int W = p.W, C = p.numCh;
  const size_t perWarp = kExactBytesPerValue * (size_t)C * W;      // (a multiple of 8)
  double *varP = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(exTerms) + warp * perWarp), *corP = varP + C * W;
  float *vals = reinterpret_cast<float *>(corP + C * W);
  for (uint32_t it = blockIdx.x * warps + warp; it < n; it += gridDim.x * warps) {
    const int64_t g = (int64_t)p.list[it];
    int lo = 0, hi = p.numFiles;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (p.fileStart[mid] <= g) lo = mid; else hi = mid;
    }
    const int64_t tl = g - p.fileStart[lo];
    const int r0 = (int)(tl % W);
    __syncwarp();
    for (int pr = 0; 2 * pr < C; pr++) {
      const float2 *src = p.data + (int64_t)pr * p.rowStride + g;
      for (int i = lane; i < W; i += 32) {
        const float2 v = src[i];
        vals[(2 * pr) * W + i] = v.x;
        if (2 * pr + 1 < C) vals[(2 * pr + 1) * W + i] = v.y;
      }
    }
    __syncwarp();
    // phase 1: the ring-order sums of the temporal group (channel 0) and the spectral group (channels 1 ..): by the whole
    // warp where the order cannot matter, else serially on lanes 0 / 1
    double sumT, sumS;
    const bool okT = ex_parallel_sum(vals, W, lane, sumT), okS = ex_parallel_sum(vals + W, (C - 1) * W, lane, sumS);
    if (!(okT && okS)) {
      double s1 = 0.0;
      if ((lane == 0 && !okT) || (lane == 1 && !okS)) s1 = ex_ring_sum(vals, W, r0, lane, lane == 0 ? 1 : C - 1);
      const double serT = __shfl_sync(0xffffffffu, s1, 0), serS = __shfl_sync(0xffffffffu, s1, 1);
      if (!okT) sumT = serT;
      if (!okS) sumS = serS;
    }
    const int nT = W, nS = (C - 1) * W;
    const double meanT = __ddiv_rn(sumT, (double)nT), meanS = __ddiv_rn(sumS, (double)nS);
    // phase 2: the TERMS of the variance (ring order, (v - mean)^2) and of the cross term (logical order, centred query x
    // (v - mean)) are independent of one another -- all lanes form them, rounded like the reference forms them, into shared
    // memory; then lanes 0 / 1 add the variance terms of T / S and lanes 2 / 3 the cross terms of T / S in the reference's
    // order: a chain step is one DADD, with nothing else between two of them
    {
      const int split = r0 == 0 ? 0 : W - r0;
      for (int e = lane; e < C * W; e += 32) {
        const int ch = e / W, k = e - ch * W;
        const double negMean = ch == 0 ? -meanT : -meanS;
        int i = k + split;
        i = i >= W ? i - W : i;
        const double xv = __dadd_rn((double)vals[ch * W + i], negMean);
        varP[e] = __dmul_rn(xv, xv);
        const double xc = __dadd_rn((double)vals[e], negMean);
        corP[e] = __dmul_rn(p.a[e], xc);
      }
    }
    __syncwarp();
    double s2 = 0.0;
    if (lane < 4) {
      const double *src = (lane < 2 ? varP : corP) + ((lane & 1) ? W : 0);
      s2 = ex_sum_doubles(src, (lane & 1) ? (C - 1) * W : W);
    }
    const double varT = __shfl_sync(0xffffffffu, s2, 0), varS = __shfl_sync(0xffffffffu, s2, 1);
    const double crT = __shfl_sync(0xffffffffu, s2, 2), crS = __shfl_sync(0xffffffffu, s2, 3);
    if (lane == 0) {
      float boost;
      if (p.boost.arr) boost = p.boost.arr[g];
      else {
        const float avg = (float)__ddiv_rn(sumT, (double)W);
        boost = (float)exp(__ddiv_rn(__dsub_rn(p.boost.lnAvgIn, log((double)avg)), 0.6));
      }
      float sim = 0.f;
      if (boost <= p.maxBoost) {
        const double stdT = __dsqrt_rn(__ddiv_rn(varT, (double)nT)), stdS = __dsqrt_rn(__ddiv_rn(varS, (double)nS));
        const float temporal = p.weight > 0.f ? (float)__ddiv_rn(crT, __dmul_rn(__dmul_rn(p.stdT, stdT), (double)nT)) : 0.f;
        const float spectral = p.weight < 1.f ? (float)__ddiv_rn(crS, __dmul_rn(__dmul_rn(p.stdS, stdS), (double)nS)) : 0.f;
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
  }
}

// boost of the offset that holds each file's maximum, by a warp per file (k_filemax_boost with the loudness window staged
// in shared memory: the ring-order Double sum is serial, its loads need not be)
__global__ void __launch_bounds__(128) k_filemax_boost_warp(BoostSrc b, const int64_t *__restrict__ fileStart,
                                                            const unsigned long long *__restrict__ fileMax, int numFiles,
                                                            float *__restrict__ out) {
  extern __shared__ float exBoostVals[];        // [warps][W]
  float *exVals = exBoostVals;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int f = blockIdx.x * warps + warp;
  if (f >= numFiles) return;
  const unsigned long long key = fileMax[f];
  if (key == 0ull) { if (lane == 0) out[f] = 1.0f; return; }
  const int64_t tl = (int64_t)(0xffffffffu - (uint32_t)key), g = fileStart[f] + tl;
  const int W = b.W;
  float *vals = exVals + (size_t)warp * W;
  for (int i = lane; i < W; i += 32) vals[i] = b.data[g + i].x;
  __syncwarp();
  double sum;
  const bool ok = ex_parallel_sum(vals, W, lane, sum);
  if (lane == 0) {
    if (!ok) sum = ex_ring_sum(vals, W, (int)(tl % W), 0, 1);
    const float avg = (float)__ddiv_rn(sum, (double)W);
    out[f] = (float)exp(__ddiv_rn(__dsub_rn(b.lnAvgIn, log((double)avg)), 0.6));
  }
}

}  // namespace sgz
