// corr_kernel.cuh -- K1: sliding-window Pearson correlation of one punch window against the
// whole database stream.
//
// Replaces hot loops A and B of the reference (FeatureCorrelationImpl.scala:190-246, :281-315):
// per frame-offset the JVM runs MathUtil.avg + 2x MathUtil.stat + MathUtil.correlate
// (MathUtil.scala:29-62,109-118,177-196) on a ring buffer.  Here:
//
//   cross(t)  = sum_c sum_i  q~[c][i] * b[c][t+i]       q~ = query - group mean (zero-mean taps)
//             = the reference's  sum (a-mean_a)(b-mean_b)  because  sum q~ = 0
//   mean_b, std_b  from window sums  S1 = sum b, S2 = sum b^2  (FP64 prefix over 14-frame chunks)
//   corr      = cross / (std_a * std_b * C * W)             (MathUtil.scala:195)
//   boost     = exp((lnAvgIn - ln(avg loudness window)) / 0.6)   (FeatureCorrelationImpl.scala:75-78)
//   sim       = boost <= maxBoost ? temporal*w + spectral*(1-w) : 0   (:199-210, Float blend)
//
// At W = 172 the path is FP32 bound (2 408 FMA per offset vs 56 B).  Measured on B200
// (tools/peaks_probe.py): scalar FFMA whose operands are live registers sustains ~51 TFLOP/s even
// in an ideal register outer product (three earlier scalar versions of this kernel all stopped at
// 41 TFLOP/s), packed FFMA2 (fma.rn.f32x2) sustains the full 73 TFLOP/s.  So the kernel is built
// around FFMA2 on CHANNEL PAIRS: the DB rows are float2 = (channel 2p, channel 2p+1), the taps are
// the matching float2, and one FFMA2 advances one offset of both channels.
//
//   * persistent CTAs, one per SM; tile = T = 14 * NC consecutive global offsets of the DB stream;
//   * warp specialisation:
//       - TMA warp: one elected lane streams the tile PAIR ROW BY PAIR ROW (one cp.async.bulk of
//         (T+Wq) float2 each) through a ring of kNSlot shared-memory slots guarded by full/empty
//         mbarriers; rows of the NEXT tile are in flight while the consumers finish the current one;
//       - stats warp: as each row lands it accumulates the per-frame spectral sums, keeps a copy of
//         the loudness channel and, after the last row, builds the FP64 chunk prefix of
//         (sum b0, sum b0^2, sum b, sum b^2) for the tile (double buffered) -- off the critical path;
//       - NC consumer threads: each owns 14 consecutive offsets for ALL channels.  Inner loop: a
//         24-frame register ring of float2 DB values + double-buffered float2 taps; per sub-step
//         (4 taps) two LDS.128 of DB values and two broadcast LDS.128 of taps, prefetched one sub-step
//         ahead straight into dead ring slots, feed 56 FFMA2 (= 112 FMA).  R = 14 gives a lane
//         stride of 112 B = 7 x 16 B: bank-conflict free for LDS.128;
//   * epilogue per offset: FP64 only where cancellation demands it (window variance), FP32 for the
//     rest; per-file maxima (first occurrence) via warp reduction + one 64-bit atomicMax.
#pragma once
#include "common.cuh"

namespace sgz {

constexpr int kR = 14;       // offsets per consumer thread
constexpr int kNSlot = 3;    // row ring depth
constexpr int kRowPad = 8;   // frames readable behind a row (prefetch overrun)
constexpr int kRing = 24;    // register ring, frames

struct CorrParams {
  const float2 *data;       // pair rows [numPairs][rowStride], normalised
  int64_t rowStride;        // frames
  int64_t usedFrames;
  int numCh;
  int numPairs;
  int W;                    // window length in frames
  int Wq;                   // W rounded up to a multiple of 4 (taps zero padded)
  const float2 *taps;       // [numPairs][Wq] zero-mean taps (group mean removed), pair interleaved
  double stdT, stdS;        // query std devs (temporal / spectral group)
  double rhoT, rhoS;        // sum of the rounded float taps per group (exact-zero correction)
  double lnAvgIn;           // ln(avg loudness of the query window)
  float weight;             // temporalWeight
  float maxBoost;
  const int64_t *fileStart; // [numFiles+1] global start frame of each file
  int numFiles;
  int tailExtra;            // frames excluded at each file end (minPunch in punch-out mode)
  int64_t tileBegin, tileEnd;   // tiles [tileBegin, tileEnd) of this launch (a streaming scan launches ranges)
  float *sim;               // [>= tileEnd*T]
  float *boost;
  unsigned long long *fileMax;  // [numFiles] packed (order_key(sim) << 32 | ~offset), or nullptr
};

struct CorrSmemLayout {
  int NC, T, rowFrames, rowPitch, numChunks;
  size_t offTaps, offStats, t0Bytes, fBytes, cpBytes, statsBytes, offBars, total;
};

// shared memory: [ring: kNSlot pair rows][taps][2 x stats{T0, F, CP, fileLoHi}][mbarriers]
__host__ __device__ inline CorrSmemLayout corr_smem_layout(int nc, int numPairs, int Wq, int nslot = kNSlot) {
  CorrSmemLayout L;
  L.NC = nc;
  L.T = kR * nc;
  L.rowFrames = L.T + Wq;
  L.rowPitch = L.rowFrames + kRowPad;
  L.numChunks = (L.rowFrames + kR - 1) / kR;
  size_t ring = (size_t)nslot * L.rowPitch * sizeof(float2);
  size_t taps = ((size_t)numPairs * Wq + kRowPad) * sizeof(float2);
  L.offTaps = ring;
  L.offStats = (L.offTaps + taps + 31) / 32 * 32;
  L.t0Bytes = ((size_t)L.rowPitch * sizeof(float) + 31) / 32 * 32;
  L.fBytes = ((size_t)L.rowPitch * sizeof(float2) + 31) / 32 * 32;
  L.cpBytes = (size_t)(L.numChunks + 1) * 4 * sizeof(double);
  L.statsBytes = (L.t0Bytes + L.fBytes + L.cpBytes + 16 + 31) / 32 * 32;
  L.offBars = L.offStats + 2 * L.statsBytes;
  L.total = L.offBars + 128;
  return L;
}

__device__ __forceinline__ float4 lds4(const void *p) { return *reinterpret_cast<const float4 *>(p); }

// ---- TMA (bulk async copy) + mbarrier primitives; SASS: UBLKCP / SYNCS ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same with an L2 eviction policy (createpolicy): data that is read once should not push reusable lines out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
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

// One sub-step = 4 taps x 14 offsets x 2 channels = 56 FFMA2.  Frame j (relative to the thread's
// first frame) lives in ring slot j % 24; sub-step s reads frames [4s, 4s+16] and meanwhile
// prefetches frames [4s+20, 4s+23] into the group of four slots that went dead with sub-step s-1.
// S = s % 6 is the compile-time phase (6 ring groups; the two tap buffers alternate with s & 1).
template <int S>
__device__ __forceinline__ void substep(float2 (&acc)[kR], float2 (&bw)[kRing], float2 (&av)[8],
                                        const float2 *__restrict__ brow, const float2 *__restrict__ arow, int s) {
  const float4 n0 = lds4(brow + 4 * s + 20), n1 = lds4(brow + 4 * s + 22);
  const float4 t0 = lds4(arow + 4 * s + 4), t1 = lds4(arow + 4 * s + 6);
  constexpr int g0 = 4 * (S % 6), a0 = 4 * (S & 1), a1 = 4 * ((S + 1) & 1), p0 = 4 * ((S + 5) % 6);
#pragma unroll
  for (int u = 0; u < 4; u++) {
#pragma unroll
    for (int r = 0; r < kR; r++) acc[r] = __ffma2_rn(av[a0 + u], bw[(g0 + u + r) % kRing], acc[r]);
  }
  bw[p0] = make_float2(n0.x, n0.y); bw[p0 + 1] = make_float2(n0.z, n0.w);
  bw[p0 + 2] = make_float2(n1.x, n1.y); bw[p0 + 3] = make_float2(n1.z, n1.w);
  av[a1] = make_float2(t0.x, t0.y); av[a1 + 1] = make_float2(t0.z, t0.w);
  av[a1 + 2] = make_float2(t1.x, t1.y); av[a1 + 3] = make_float2(t1.z, t1.w);
}

// 14 offsets x Wq taps of one channel pair; brow = &row[o], arow = &taps[p][0]; nSub = Wq / 4.
// The prefetches of the sub-steps after the last one read <= kRowPad frames past the row (padding).
__device__ __forceinline__ void conv_pair(float2 (&acc)[kR], const float2 *__restrict__ brow,
                                          const float2 *__restrict__ arow, int nSub) {
  float2 bw[kRing], av[8];
#pragma unroll
  for (int g = 0; g < 5; g++) {
    const float4 v0 = lds4(brow + 4 * g), v1 = lds4(brow + 4 * g + 2);
    bw[4 * g] = make_float2(v0.x, v0.y); bw[4 * g + 1] = make_float2(v0.z, v0.w);
    bw[4 * g + 2] = make_float2(v1.x, v1.y); bw[4 * g + 3] = make_float2(v1.z, v1.w);
  }
#pragma unroll
  for (int k = 20; k < kRing; k++) bw[k] = make_float2(0.f, 0.f);
  {
    const float4 t0 = lds4(arow), t1 = lds4(arow + 2);
    av[0] = make_float2(t0.x, t0.y); av[1] = make_float2(t0.z, t0.w);
    av[2] = make_float2(t1.x, t1.y); av[3] = make_float2(t1.z, t1.w);
#pragma unroll
    for (int k = 4; k < 8; k++) av[k] = make_float2(0.f, 0.f);
  }
  int s = 0;
#pragma unroll 1
  for (; s + 6 <= nSub; s += 6) {
    substep<0>(acc, bw, av, brow, arow, s);
    substep<1>(acc, bw, av, brow, arow, s + 1);
    substep<2>(acc, bw, av, brow, arow, s + 2);
    substep<3>(acc, bw, av, brow, arow, s + 3);
    substep<4>(acc, bw, av, brow, arow, s + 4);
    substep<5>(acc, bw, av, brow, arow, s + 5);
  }
  const int rem = nSub - s;   // 0..5, warp uniform
  if (rem > 0) substep<0>(acc, bw, av, brow, arow, s);
  if (rem > 1) substep<1>(acc, bw, av, brow, arow, s + 1);
  if (rem > 2) substep<2>(acc, bw, av, brow, arow, s + 2);
  if (rem > 3) substep<3>(acc, bw, av, brow, arow, s + 3);
  if (rem > 4) substep<4>(acc, bw, av, brow, arow, s + 4);
}

struct D4 {
  double t1, t2, s1, s2;
};

// blockDim.x = NC + 64: warps [0, NC/32) consumers, warp NC/32 = TMA issuer, warp NC/32+1 = stats.
template <int NSLOT>
__global__ void __launch_bounds__(384, 1) k_corr(const CorrParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int NC = (int)blockDim.x - 64;
  const int NCW = NC >> 5;
  const int Wq = p.Wq, W = p.W;
  const CorrSmemLayout L = corr_smem_layout(NC, p.numPairs, Wq, NSLOT);
  const int T = L.T, rowFrames = L.rowFrames, rowPitch = L.rowPitch, numChunks = L.numChunks;

  float2 *ring = reinterpret_cast<float2 *>(smem);
  float2 *taps = reinterpret_cast<float2 *>(smem + L.offTaps);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.offBars);
  uint64_t *full = bars, *empty = bars + NSLOT, *statsReady = bars + 2 * NSLOT, *statsFree = bars + 2 * NSLOT + 2;
  auto stats_T0 = [&](int b) { return reinterpret_cast<float *>(smem + L.offStats + b * L.statsBytes); };
  auto stats_F = [&](int b) { return reinterpret_cast<float2 *>(smem + L.offStats + b * L.statsBytes + L.t0Bytes); };
  auto stats_CP = [&](int b) {
    return reinterpret_cast<double *>(smem + L.offStats + b * L.statsBytes + L.t0Bytes + L.fBytes);
  };
  auto stats_file = [&](int b) {
    return reinterpret_cast<int *>(smem + L.offStats + b * L.statsBytes + L.t0Bytes + L.fBytes + L.cpBytes);
  };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NCW + 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(statsReady + b, 1);
      mbar_init(statsFree + b, NCW);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < p.numPairs * Wq + kRowPad; i += blockDim.x)
    taps[i] = i < p.numPairs * Wq ? p.taps[i] : make_float2(0.f, 0.f);
  __syncthreads();

  const uint32_t rowBytes = (uint32_t)rowFrames * sizeof(float2);

  if (warp == NCW) {
    // =========================== TMA issuer ===========================
    if (lane == 0) {
      uint32_t rc = 0;  // rows issued so far by this CTA
      for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x) {
        const int64_t t0 = tile * T;
        for (int c = 0; c < p.numPairs; c++, rc++) {
          const int slot = rc % NSLOT;
          mbar_wait(empty + slot, ((rc / NSLOT) & 1) ^ 1);
          mbar_expect_tx(full + slot, rowBytes);
          bulk_g2s(ring + (size_t)slot * rowPitch, p.data + (int64_t)c * p.rowStride + t0, rowBytes, full + slot);
        }
      }
    }
  } else if (warp == NCW + 1) {
    // =========================== stats warp ===========================
    uint32_t rc = 0, it = 0;
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, it++) {
      const int b = it & 1;
      float *T0 = stats_T0(b);
      float2 *F = stats_F(b);
      double *CP = stats_CP(b);
      mbar_wait(statsFree + b, ((it >> 1) & 1) ^ 1);   // consumers are done with this buffer (tile it-2)
      const int64_t t0 = tile * T;
      if (lane < 2) {  // file range of the tile
        int64_t g = lane == 0 ? t0 : min(t0 + T - 1, p.usedFrames - 1);
        int lo = 0, hi = p.numFiles;
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (p.fileStart[mid] <= g) lo = mid; else hi = mid;
        }
        stats_file(b)[lane] = lo;
      }
      const int nv = rowFrames >> 1;  // two frames (one LDS.128) per iteration; rowFrames is even
      for (int c = 0; c < p.numPairs; c++, rc++) {
        const int slot = rc % NSLOT;
        mbar_wait(full + slot, (rc / NSLOT) & 1);
        const float2 *row = ring + (size_t)slot * rowPitch;
        if (c == 0) {   // pair 0 = (loudness, first spectral channel)
          for (int v = lane; v < nv; v += 32) {
            const float4 x = lds4(row + 2 * v);
            *reinterpret_cast<float2 *>(T0 + 2 * v) = make_float2(x.x, x.z);
            *reinterpret_cast<float4 *>(F + 2 * v) = make_float4(x.y, x.y * x.y, x.w, x.w * x.w);
          }
        } else {
#pragma unroll 4
          for (int v = lane; v < nv; v += 32) {
            const float4 x = lds4(row + 2 * v);
            float4 f = *reinterpret_cast<float4 *>(F + 2 * v);
            f.x += x.x + x.y; f.y = fmaf(x.x, x.x, fmaf(x.y, x.y, f.y));
            f.z += x.z + x.w; f.w = fmaf(x.z, x.z, fmaf(x.w, x.w, f.w));
            *reinterpret_cast<float4 *>(F + 2 * v) = f;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
      }
      // FP64 prefix over 14-frame chunks: CP[j] = sum over frames < 14 j of (b0, b0^2, sum_c b, sum_c b^2)
      const int per = (numChunks + 31) / 32;
      const int jb = min(lane * per, numChunks), je = min(jb + per, numChunks);
      double r1 = 0, r2 = 0, r3 = 0, r4 = 0;
      for (int j = jb; j < je; j++) {
        double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        const int e0 = kR * j, e1 = min(kR * (j + 1), rowFrames);
        if (e1 - e0 == kR) {
          // full chunk: all loads first, then a pairwise tree (depth 4 instead of a 14-long FP64 chain)
          double d1[kR], d2[kR], d3[kR], d4[kR];
#pragma unroll
          for (int k = 0; k < kR; k += 2) {
            const float2 t = *reinterpret_cast<const float2 *>(T0 + e0 + k);
            const float4 f = *reinterpret_cast<const float4 *>(F + e0 + k);
            d1[k] = (double)t.x; d1[k + 1] = (double)t.y;
            d3[k] = (double)f.x; d4[k] = (double)f.y; d3[k + 1] = (double)f.z; d4[k + 1] = (double)f.w;
          }
#pragma unroll
          for (int k = 0; k < kR; k++) d2[k] = d1[k] * d1[k];
#pragma unroll
          for (int w = 1; w < kR; w <<= 1) {
#pragma unroll
            for (int k = 0; k + w < kR; k += 2 * w) {
              d1[k] += d1[k + w]; d2[k] += d2[k + w]; d3[k] += d3[k + w]; d4[k] += d4[k + w];
            }
          }
          a1 = d1[0]; a2 = d2[0]; a3 = d3[0]; a4 = d4[0];
        } else {
          for (int e = e0; e < e1; e++) {
            const double b0 = (double)T0[e];
            const float2 f = F[e];
            a1 += b0; a2 += b0 * b0; a3 += (double)f.x; a4 += (double)f.y;
          }
        }
        r1 += a1; r2 += a2; r3 += a3; r4 += a4;
        double *o = CP + 4 * (size_t)(j + 1);
        o[0] = r1; o[1] = r2; o[2] = r3; o[3] = r4;
      }
      double i1 = r1, i2 = r2, i3 = r3, i4 = r4;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o1 = __shfl_up_sync(0xffffffffu, i1, d), o2 = __shfl_up_sync(0xffffffffu, i2, d);
        const double o3 = __shfl_up_sync(0xffffffffu, i3, d), o4 = __shfl_up_sync(0xffffffffu, i4, d);
        if (lane >= d) { i1 += o1; i2 += o2; i3 += o3; i4 += o4; }
      }
      const double x1 = i1 - r1, x2 = i2 - r2, x3 = i3 - r3, x4 = i4 - r4;
      for (int j = jb; j < je; j++) {
        double *o = CP + 4 * (size_t)(j + 1);
        o[0] += x1; o[1] += x2; o[2] += x3; o[3] += x4;
      }
      if (lane == 0) { CP[0] = 0; CP[1] = 0; CP[2] = 0; CP[3] = 0; }
      __syncwarp();
      if (lane == 0) mbar_arrive(statsReady + b);
    }
  } else {
    // =========================== consumers ===========================
    const int o = tid * kR;
    const int nSub = Wq >> 2;
    const double invW = 1.0 / (double)W, invNS = 1.0 / ((double)(p.numCh - 1) * (double)W);
    const float cT = (float)(invW / p.stdT), cS = (float)(invNS / p.stdS);
    const float rhoT = (float)p.rhoT, rhoS = (float)p.rhoS, l2In = (float)(p.lnAvgIn * 1.4426950408889634);
    const bool useT = p.weight > 0.f, useS = p.weight < 1.f;
    const float qnan = __int_as_float(0x7fc00000);
    const int nq = W / kR, remW = W - nq * kR;
    uint32_t rc = 0, it = 0;
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, it++) {
      const int64_t t0 = tile * T;
      float2 acc[kR];
      float accT[kR];
#pragma unroll
      for (int r = 0; r < kR; r++) acc[r] = make_float2(0.f, 0.f);
      for (int c = 0; c < p.numPairs; c++, rc++) {
        const int slot = rc % NSLOT;
        mbar_wait(full + slot, (rc / NSLOT) & 1);
        conv_pair(acc, ring + (size_t)slot * rowPitch + o, taps + c * Wq, nSub);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        if (c == 0) {   // .x of pair 0 is the temporal (loudness) cross term
#pragma unroll
          for (int r = 0; r < kR; r++) { accT[r] = acc[r].x; acc[r].x = 0.f; }
        }
      }
      // ---- epilogue ----
      const int b = it & 1;
      mbar_wait(statsReady + b, (it >> 1) & 1);
      const float *T0 = stats_T0(b);
      const float2 *F = stats_F(b);
      const double *CP = stats_CP(b);
      const int fLo = stats_file(b)[0], fHi = stats_file(b)[1];
      D4 win;
      {
        const double *c0 = CP + 4 * (size_t)tid, *c1 = CP + 4 * (size_t)(tid + nq);
        win = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2], c1[3] - c0[3]};
        for (int k = 0; k < remW; k++) {
          const int e = o + kR * nq + k;
          const double b0 = (double)T0[e];
          const float2 f = F[e];
          win.t1 += b0; win.t2 += b0 * b0; win.s1 += (double)f.x; win.s2 += (double)f.y;
        }
      }
      const int64_t g0 = t0 + o;
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
      // pass A: straight-line arithmetic for all 14 offsets (no file logic in here, so the 14 chains overlap).
      // FP64 only for the window variance; boost via MUFU lg2/ex2 (|rel err| < 1e-6); rsqrt instead of sqrt+div.
      float simv[kR], boostv[kR];
#pragma unroll
      for (int k = 0; k < kR; k++) {
        const double mT = win.t1 * invW;
        const float avgB = (float)mT;                                          // MathUtil.avg -> Float
        const float boost = exp2f((l2In - __log2f(avgB)) * (1.0f / 0.6f));     // calcBoost
        float temporal = 0.f, spectral = 0.f;
        if (useT) {
          const double q = win.t2 * invW;
          const double var = q - mT * mT;
          const float cr = accT[k] - (float)mT * rhoT;
          temporal = (var > 1e-13 * q) ? (cr * cT) * rsqrtf((float)var) : qnan;
        }
        if (useS) {
          const double mS = win.s1 * invNS;
          const double q = win.s2 * invNS;
          const double var = q - mS * mS;
          const float cr = (acc[k].x + acc[k].y) - (float)mS * rhoS;
          spectral = (var > 1e-13 * q) ? (cr * cS) * rsqrtf((float)var) : qnan;
        }
        const float blend = __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
        simv[k] = boost <= p.maxBoost ? blend : 0.f;
        boostv[k] = boost;
        if (k < kR - 1) {  // slide the window by one frame
          const int e = o + k;
          const double bo = (double)T0[e], bn = (double)T0[e + W];
          const float2 fo = F[e], fn = F[e + W];
          win.t1 += bn - bo;
          win.t2 += bn * bn - bo * bo;
          win.s1 += (double)fn.x - (double)fo.x;
          win.s2 += (double)fn.y - (double)fo.y;
        }
      }
      // pass B: which offsets exist (window inside its file), per-file maximum (first occurrence)
      unsigned long long best = 0ull;
      int bestFile = -1;
      bool straddle = false;
      if (g0 + kR <= fEnd || f + 1 >= p.numFiles) {
        // common case: all 14 offsets belong to one file
        const int64_t lim = fEnd - p.tailExtra - W + 1 - g0;          // offsets k < lim are evaluated
        const int nOk = lim < 0 ? 0 : (lim > kR ? kR : (int)lim);
        const uint32_t tl0 = (uint32_t)(g0 - fStart);
#pragma unroll
        for (int k = 0; k < kR; k++) {
          if (k >= nOk) { simv[k] = qnan; boostv[k] = qnan; }
          else if (simv[k] == simv[k]) {
            const unsigned long long key = ((unsigned long long)float_order_key(simv[k]) << 32) |
                                           (unsigned long long)(0xffffffffu - (tl0 + (uint32_t)k));
            if (key > best) { best = key; bestFile = f; }
          }
        }
      } else {
        for (int k = 0; k < kR; k++) {
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
          float sv = qnan, bv = qnan;
#pragma unroll
          for (int kk = 0; kk < kR; kk++) if (kk == k) { sv = simv[kk]; bv = boostv[kk]; }
          if (!(g < p.usedFrames && tl < nValid)) { sv = qnan; bv = qnan; }
          else if (sv == sv) {
            const unsigned long long key = ((unsigned long long)float_order_key(sv) << 32) |
                                           (unsigned long long)(0xffffffffu - (uint32_t)tl);
            if (key > best) { best = key; bestFile = f; }
          }
#pragma unroll
          for (int kk = 0; kk < kR; kk++) if (kk == k) { simv[kk] = sv; boostv[kk] = bv; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(statsFree + b);   // stats buffer b may be rebuilt (tile it+2)
      {
        float2 *so = reinterpret_cast<float2 *>(p.sim + g0);    // g0 is even: 8-byte aligned
        float2 *bo = reinterpret_cast<float2 *>(p.boost + g0);
#pragma unroll
        for (int k = 0; k < kR / 2; k++) {
          so[k] = make_float2(simv[2 * k], simv[2 * k + 1]);
          bo[k] = make_float2(boostv[2 * k], boostv[2 * k + 1]);
        }
      }
      if (p.fileMax) {
        const unsigned fullm = 0xffffffffu;
        const int f0 = __reduce_max_sync(fullm, bestFile);
        const bool uniform = __all_sync(fullm, (bestFile == f0 || best == 0ull) && !straddle) && f0 >= 0;
        if (uniform) {
          unsigned long long m = best;
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o2 = __shfl_xor_sync(fullm, m, d);
            m = o2 > m ? o2 : m;
          }
          if (lane == 0 && m != 0ull) atomicMax(p.fileMax + f0, m);
        } else if (best != 0ull) {
          atomicMax(p.fileMax + bestFile, best);
        }
      }
    }
  }
}

}  // namespace sgz
