// corr_kernel.cuh -- K1: sliding-window Pearson correlation of one punch window against the
// whole database stream.
//
// Replaces hot loops A and B of the reference (FeatureCorrelationImpl.scala:190-246, :281-315):
// per frame-offset the JVM runs MathUtil.avg + 2x MathUtil.stat + MathUtil.correlate
// (MathUtil.scala:29-62,109-118,177-196) on a ring buffer.  Here:
//
//   cross(t)  = sum_c sum_i  q~[c][i] * b[c][t+i]       q~ = query - group mean (zero-mean taps)
//             = the reference's  sum (a-mean_a)(b-mean_b)  because  sum q~ = 0
//   mean_b, std_b  from window sums  S1 = sum b, S2 = sum b^2  (FP64), obtained from per-frame
//             sums + an FP64 chunk prefix inside the tile (no second pass over the window)
//   corr      = cross / (std_a * std_b * C * W)             (MathUtil.scala:195)
//   boost     = exp((lnAvgIn - ln(avg loudness window)) / 0.6)   (FeatureCorrelationImpl.scala:75-78)
//   sim       = boost <= maxBoost ? temporal*w + spectral*(1-w) : 0   (:199-210, Float blend)
//
// Work decomposition (B200: 148 SMs, 227 KB smem/SM, 128 FP32 lanes/SM/clk):
//   * the DB is one planar stream data[c][g]; a CTA owns T = 12*NTG consecutive global offsets
//     and stages frames [t0, t0+T+Wq) of all channels in shared memory once (56 B/frame of HBM
//     traffic, halo (W-1)/T);
//   * every thread owns R = 12 consecutive offsets and keeps a 16-register sliding window of
//     database values, so one LDS.128 of DB data + one broadcast LDS.128 of taps feeds 48 FFMA;
//     R = 12 makes the lane stride 48 B, which is bank-conflict free for LDS.128 without padding;
//   * the channels are split over two thread groups (2*NTG threads per CTA) so that 8 warps per
//     SM hide shared-memory latency; partial sums are exchanged through shared memory once per
//     tile and each group finishes 6 of the 12 offsets.
//   * per-file maxima (first occurrence) are reduced per warp and merged with one 64-bit
//     atomicMax -- they drive the candidate filter of K2 (select.cuh).
#pragma once
#include "common.cuh"

namespace sgz {

constexpr int kR = 12;  // offsets per thread

struct CorrParams {
  const float *data;        // planar DB [numCh][chanStride], normalised
  int64_t chanStride;
  int64_t usedFrames;
  int numCh;
  int csplit;               // channels [0,csplit) -> group 0, [csplit,numCh) -> group 1
  int W;                    // window length in frames
  int Wq;                   // W rounded up to a multiple of 4 (taps zero padded)
  const float *taps;        // [numCh][Wq] zero-mean taps (group mean removed)
  double stdT, stdS;        // query std devs (temporal / spectral group)
  double rhoT, rhoS;        // sum of the rounded float taps per group (exact-zero correction)
  double lnAvgIn;           // ln(avg loudness of the query window)
  float weight;             // temporalWeight
  float maxBoost;
  const int64_t *fileStart; // [numFiles+1] global start frame of each file
  int numFiles;
  int tailExtra;            // frames excluded at each file end (minPunch in punch-out mode)
  float *sim;               // [>= numTiles*T]
  float *boost;
  unsigned long long *fileMax;  // [numFiles] packed (order_key(sim) << 32 | ~offset), or nullptr
};

struct CorrSmemLayout {
  int T, tileFrames, pitch, numChunks;
  size_t offTaps, offF, offCP, offFile, total;
};

inline CorrSmemLayout corr_smem_layout(int ntg, int numCh, int Wq) {
  CorrSmemLayout L;
  L.T = kR * ntg;
  L.tileFrames = L.T + Wq;
  L.pitch = L.tileFrames;
  L.numChunks = L.tileFrames / 4;
  size_t tile = (size_t)numCh * L.pitch * sizeof(float);
  size_t taps = std::max((size_t)numCh * Wq, (size_t)18 * ntg) * sizeof(float);
  size_t F = (size_t)L.tileFrames * sizeof(float2);
  size_t CP = (size_t)(L.numChunks + 1) * 4 * sizeof(double);
  L.offTaps = tile;
  L.offF = (L.offTaps + taps + 15) / 16 * 16;
  L.offCP = (L.offF + F + 31) / 32 * 32;
  L.offFile = L.offCP + CP;
  L.total = L.offFile + 32;
  return L;
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// ---- TMA (bulk async copy) + mbarrier primitives; SASS: UBLKCP / SYNCS ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_LOOP_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One sub-step = 4 taps x 12 offsets = 48 FFMA fed by one LDS.128 of DB values and one broadcast
// LDS.128 of taps, both prefetched one sub-step ahead (software pipelining: with 2-3 warps per
// scheduler the 30-cycle LDS latency must not sit between the load and its first FFMA).
#define SGZ_SUBSTEP(S, NEXT)                                                   \
  {                                                                            \
    bw[(12 + 4 * (S)) & 15] = nb.x;                                            \
    bw[(13 + 4 * (S)) & 15] = nb.y;                                            \
    bw[(14 + 4 * (S)) & 15] = nb.z;                                            \
    bw[(15 + 4 * (S)) & 15] = nb.w;                                            \
    const float av0 = a.x, av1 = a.y, av2 = a.z, av3 = a.w;                    \
    nb = lds4(brow + 4 * (NEXT) + 12);                                         \
    a = lds4(arow + 4 * (NEXT));                                               \
    _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = fmaf(av0, bw[(4 * (S) + 0 + r) & 15], acc[r]); \
    _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = fmaf(av1, bw[(4 * (S) + 1 + r) & 15], acc[r]); \
    _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = fmaf(av2, bw[(4 * (S) + 2 + r) & 15], acc[r]); \
    _Pragma("unroll") for (int r = 0; r < kR; r++) acc[r] = fmaf(av3, bw[(4 * (S) + 3 + r) & 15], acc[r]); \
  }

// 12 offsets x Wq taps of one channel; brow = &tile[c][o], arow = &taps[c][0]; nSub = Wq / 4.
// The prefetch of the sub-step after the last one reads <= 16 B past the row (inside the smem
// allocation, value unused).
__device__ __forceinline__ void conv_channel(float (&acc)[kR], const float *__restrict__ brow,
                                             const float *__restrict__ arow, int nSub) {
  float bw[16];
  {
    float4 v0 = lds4(brow), v1 = lds4(brow + 4), v2 = lds4(brow + 8);
    bw[0] = v0.x; bw[1] = v0.y; bw[2] = v0.z; bw[3] = v0.w;
    bw[4] = v1.x; bw[5] = v1.y; bw[6] = v1.z; bw[7] = v1.w;
    bw[8] = v2.x; bw[9] = v2.y; bw[10] = v2.z; bw[11] = v2.w;
  }
  float4 nb = lds4(brow + 12), a = lds4(arow);
  int sub = 0;
#pragma unroll 1
  for (; sub + 4 <= nSub; sub += 4) {
    SGZ_SUBSTEP(0, sub + 1)
    SGZ_SUBSTEP(1, sub + 2)
    SGZ_SUBSTEP(2, sub + 3)
    SGZ_SUBSTEP(3, sub + 4)
  }
  if (sub < nSub) {
    SGZ_SUBSTEP(0, sub + 1)
    if (sub + 1 < nSub) {
      SGZ_SUBSTEP(1, sub + 2)
      if (sub + 2 < nSub) { SGZ_SUBSTEP(2, sub + 3) }
    }
  }
}

struct D4 {
  double t1, t2, s1, s2;
};

template <int NTG>
__global__ void __launch_bounds__(2 * NTG, 2) k_corr(const CorrParams p) {
  constexpr int T = kR * NTG;
  constexpr int NT = 2 * NTG;
  extern __shared__ __align__(128) unsigned char smem[];
  const int Wq = p.Wq;
  const int tileFrames = T + Wq;
  const int pitch = tileFrames;
  const int numChunks = tileFrames >> 2;

  float *tile = reinterpret_cast<float *>(smem);
  size_t offTaps = (size_t)p.numCh * pitch * sizeof(float);
  size_t tapsBytes = (size_t)max(p.numCh * Wq, 18 * NTG) * sizeof(float);
  size_t offF = (offTaps + tapsBytes + 15) / 16 * 16;
  size_t offCP = (offF + (size_t)tileFrames * sizeof(float2) + 31) / 32 * 32;
  size_t offFile = offCP + (size_t)(numChunks + 1) * 4 * sizeof(double);
  float *taps = reinterpret_cast<float *>(smem + offTaps);
  float2 *F = reinterpret_cast<float2 *>(smem + offF);
  double *CP = reinterpret_cast<double *>(smem + offCP);      // [numChunks+1][4]
  int *shFile = reinterpret_cast<int *>(smem + offFile);      // [0]=file of t0, [1]=file of last frame
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + offFile + 8);

  const int tid = threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.x * T;

  // ---- stage taps + tile with TMA bulk copies: one elected thread, 1 + numCh copies, one mbarrier ----
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    const uint32_t rowBytes = (uint32_t)tileFrames * sizeof(float);
    const uint32_t tapBytes = (uint32_t)(p.numCh * Wq) * sizeof(float);
    mbar_expect_tx(bar, rowBytes * p.numCh + tapBytes);
    bulk_g2s(taps, p.taps, tapBytes, bar);
    for (int c = 0; c < p.numCh; c++) bulk_g2s(tile + c * pitch, p.data + (int64_t)c * p.chanStride + t0, rowBytes, bar);
  }
  // ---- file range of this tile (two threads, overlaps with the copies) ----
  if (tid >= 32 && tid < 34) {
    int64_t g = tid == 32 ? t0 : min(t0 + T - 1, p.usedFrames - 1);
    int lo = 0, hi = p.numFiles;  // f with fileStart[f] <= g < fileStart[f+1]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.fileStart[mid] <= g) lo = mid; else hi = mid;
    }
    shFile[tid - 32] = lo;
  }
  if (tid < 4) CP[tid] = 0.0;
  __syncthreads();          // barrier init visible to the waiters
  mbar_wait(bar, 0);

  // ---- per-frame spectral sums (float) ----
  for (int e = tid; e < tileFrames; e += NT) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = 1; c < p.numCh; c++) {
      float b = tile[c * pitch + e];
      s1 += b;
      s2 = fmaf(b, b, s2);
    }
    F[e] = make_float2(s1, s2);
  }
  __syncthreads();
  // ---- FP64 chunk sums (4 frames per chunk), then an exclusive scan: CP[j] = sum of chunks < j ----
  for (int j = tid; j < numChunks; j += NT) {
    double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double b0 = (double)tile[4 * j + k];
      float2 f = F[4 * j + k];
      a1 += b0;
      a2 += b0 * b0;
      a3 += (double)f.x;
      a4 += (double)f.y;
    }
    double *o = CP + 4 * (size_t)(j + 1);
    o[0] = a1; o[1] = a2; o[2] = a3; o[3] = a4;
  }
  __syncthreads();
  {
    // warp w scans component w: lane l owns entries [l*per, (l+1)*per) of CP[1..numChunks]
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 4) {
      const int per = (numChunks + 31) / 32;
      const int b = lane * per, e = min(b + per, numChunks);
      double run = 0.0;
      for (int j = b; j < e; j++) {
        run += CP[4 * (size_t)(j + 1) + warp];
        CP[4 * (size_t)(j + 1) + warp] = run;
      }
      double incl = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      double excl = incl - run;
      for (int j = b; j < e; j++) CP[4 * (size_t)(j + 1) + warp] += excl;
    }
  }
  // (the main loop below only reads tile/taps; CP/F are consumed after the next barrier)

  // ---- main loop: FFMA sliding correlation ----
  const int grp = tid / NTG;       // warp-uniform (NTG % 32 == 0)
  const int lt = tid - grp * NTG;
  const int o = lt * kR;
  const int nSub = Wq >> 2;
  float accT[kR], accS[kR];
#pragma unroll
  for (int r = 0; r < kR; r++) { accT[r] = 0.f; accS[r] = 0.f; }
  if (grp == 0) {
    conv_channel(accT, tile + o, taps, nSub);
    for (int c = 1; c < p.csplit; c++) conv_channel(accS, tile + c * pitch + o, taps + c * Wq, nSub);
  } else {
    for (int c = p.csplit; c < p.numCh; c++) conv_channel(accS, tile + c * pitch + o, taps + c * Wq, nSub);
  }
  __syncthreads();  // everybody is done with the taps; CP is complete

  // ---- exchange partial sums: group g finishes offsets r in [6g, 6g+6) ----
  float *xch = taps;  // aliases the taps region: [18][NTG]
  if (grp == 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      xch[(6 + k) * NTG + lt] = accT[6 + k];
      xch[(12 + k) * NTG + lt] = accS[6 + k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 6; k++) xch[k * NTG + lt] = accS[k];
  }
  __syncthreads();
  float crossT[6], crossS[6];
  if (grp == 0) {
#pragma unroll
    for (int k = 0; k < 6; k++) { crossT[k] = accT[k]; crossS[k] = accS[k] + xch[k * NTG + lt]; }
  } else {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      crossT[k] = xch[(6 + k) * NTG + lt];
      crossS[k] = accS[6 + k] + xch[(12 + k) * NTG + lt];
    }
  }

  // ---- window sums (FP64): start chunk aligned at o + 4*grp, slide to o + 6*grp ----
  const int W = p.W;
  const int nq = W >> 2, rem = W & 3;
  int ws = o + 4 * grp;  // window start (tile-local frame), multiple of 4
  D4 win;
  {
    const double *c0 = CP + 4 * (size_t)(ws >> 2), *c1 = CP + 4 * (size_t)((ws >> 2) + nq);
    win = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2], c1[3] - c0[3]};
    for (int k = 0; k < rem; k++) {
      int e = ws + 4 * nq + k;
      double b0 = (double)tile[e];
      float2 f = F[e];
      win.t1 += b0; win.t2 += b0 * b0; win.s1 += (double)f.x; win.s2 += (double)f.y;
    }
  }
  auto slide = [&](int start) {  // window [start, start+W) -> [start+1, start+1+W)
    double bo = (double)tile[start], bn = (double)tile[start + W];
    float2 fo = F[start], fn = F[start + W];
    win.t1 += bn - bo;
    win.t2 += bn * bn - bo * bo;
    win.s1 += (double)fn.x - (double)fo.x;
    win.s2 += (double)fn.y - (double)fo.y;
  };
  if (grp == 1) { slide(ws); slide(ws + 1); ws += 2; }

  // ---- epilogue: FP64 only where cancellation demands it (variance), FP32 elsewhere ----
  const int fLo = shFile[0], fHi = shFile[1];
  const int64_t g0 = t0 + ws;
  int f = fLo;
  {
    int lo = fLo, hi = fHi + 1;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (p.fileStart[mid] <= g0) lo = mid; else hi = mid;
    }
    f = lo;
  }
  int64_t fStart = p.fileStart[f], fEnd = p.fileStart[f + 1];
  const double invW = 1.0 / (double)W, invNS = 1.0 / ((double)(p.numCh - 1) * (double)W);
  const float cT = (float)(invW / p.stdT), cS = (float)(invNS / p.stdS);
  const float rhoT = (float)p.rhoT, rhoS = (float)p.rhoS, lnIn = (float)p.lnAvgIn;
  const bool useT = p.weight > 0.f, useS = p.weight < 1.f;
  const float qnan = __int_as_float(0x7fc00000);
  float simv[6], boostv[6];
  unsigned long long best = 0ull;
  int bestFile = -1;
  bool straddle = false;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const int64_t g = g0 + k;
    while (g >= fEnd && f + 1 < p.numFiles) {
      if (best != 0ull && p.fileMax) { atomicMax(p.fileMax + bestFile, best); straddle = true; }
      best = 0ull;
      f++;
      fStart = fEnd;
      fEnd = p.fileStart[f + 1];
    }
    const int64_t tl = g - fStart;
    const int64_t nValid = (fEnd - fStart) - p.tailExtra - W + 1;
    float sim = qnan, boost = qnan;
    if (g < p.usedFrames && tl < nValid) {
      const double mT = win.t1 * invW;
      const float avgB = (float)mT;                              // MathUtil.avg -> Float
      boost = expf((lnIn - logf(avgB)) / 0.6f);                  // calcBoost
      if (boost <= p.maxBoost) {
        float temporal = 0.f, spectral = 0.f;
        if (useT) {
          const double q = win.t2 * invW;
          const double var = q - mT * mT;
          const float cr = crossT[k] - (float)mT * rhoT;
          temporal = (var > 1e-13 * q) ? (cr * cT) / sqrtf((float)var) : qnan;
        }
        if (useS) {
          const double mS = win.s1 * invNS;
          const double q = win.s2 * invNS;
          const double var = q - mS * mS;
          const float cr = crossS[k] - (float)mS * rhoS;
          spectral = (var > 1e-13 * q) ? (cr * cS) / sqrtf((float)var) : qnan;
        }
        sim = __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
      } else {
        sim = 0.f;
      }
      if (sim == sim) {
        unsigned long long key = ((unsigned long long)float_order_key(sim) << 32) |
                                 (unsigned long long)(0xffffffffu - (uint32_t)tl);
        if (key > best) { best = key; bestFile = f; }
      }
    }
    simv[k] = sim;
    boostv[k] = boost;
    if (k < 5) slide(ws + k);
  }
  // 6 contiguous floats, 8-byte aligned
  {
    float2 *so = reinterpret_cast<float2 *>(p.sim + g0);
    float2 *bo = reinterpret_cast<float2 *>(p.boost + g0);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      so[k] = make_float2(simv[2 * k], simv[2 * k + 1]);
      bo[k] = make_float2(boostv[2 * k], boostv[2 * k + 1]);
    }
  }
  if (p.fileMax) {
    // warp-aggregate when every lane that found something sits in the same file
    const unsigned full = 0xffffffffu;
    const int f0 = __reduce_max_sync(full, bestFile);
    const bool uniform = __all_sync(full, (bestFile == f0 || best == 0ull) && !straddle) && f0 >= 0;
    if (uniform) {
      unsigned long long m = best;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o2 = __shfl_xor_sync(full, m, d);
        m = o2 > m ? o2 : m;
      }
      if ((tid & 31) == 0 && m != 0ull) atomicMax(p.fileMax + f0, m);
    } else if (best != 0ull) {
      atomicMax(p.fileMax + bestFile, best);
    }
  }
}

}  // namespace sgz
