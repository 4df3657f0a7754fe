// selfsim_tc.cuh -- K4 on the tensor cores (tcgen05 / TMEM): the SelfSimilarity Gram matrix of selfsim_fast.cuh with
// split-FP16 operands.
//
//   G[a][b] = sum_c sum_h x1[c][d a + h] * x2[c][d b + h]      (closed form of MathUtil.correlateHalf, SURVEY.md 3.3;
//                                                               SelfSimilarityImpl.scala:127-155)
// per 128 x 128 cell tile of the upper triangle: M = 128 windows of file 1, N = 128 windows of file 2, K = channels x H.
// Precision: the centred data are scaled by a power of two into the FP16 range and split into two FP16 numbers
// (22 significant bits), a*b = a1*b1 + (a2*b1 + a1*b2).  The tensor core truncates when it aligns addends, so the number
// of accumulations into a LARGE accumulator bounds the bias: the main products of the spectral group get their own TMEM
// region (6 x 13 = 78 MMAs at H = 86), the small correction products another, the temporal group (18 MMAs) a third.
//
// (Measured worst deviation from the oracle 3.8e-6 relative at H = 86; the host routes windows whose spectral chain would
// exceed 78 MMAs to the FFMA2 kernel, see api_segself.cuh and tools/selfsim_error_probe.py.)
//
// Operands are windows of ONE signal (Hankel matrices), so nothing is materialised in global memory except a "record"
// array R[part][c][rho] = the 8 FP16 values of frames g rho .. g rho + 7 (g = gcd(decim, 8)): any 16-byte chunk
// (window a, k = 8 kc .. 8 kc + 7) of an operand tile is then ONE aligned record, rho = (decim / g) a + (8 / g) kc.
//  * In-place mode (decim | 8): a window starts at every record, so the 8 rows of a no-swizzle K-major core matrix ARE 8
//    consecutive records and the record stage in shared memory is itself the operand of both sides (SBO = 128 B,
//    LBO = (8 / g) * 16 B, overlapping core matrices).  Only B's last K step, which must end at the window's end, comes
//    pre-masked from global memory (k_self_tail) as part of the stage.  Warp 0 streams the stages with bulk copies
//    through a ring of 5 - 8; nobody touches the operands with generic loads / stores.
//  * Expansion mode (other decimations): 6 builder warps expand the records of a stage into canonical core matrices
//    (one LDS.128 + one STS.128 per chunk, conflict free), zeroing k >= H, in a ring of slabs of a few K steps.
// One issuer warp (elected lane) feeds tcgen05.mma M128 x N128 x K16 (kind::f16, FP32 accumulate; per K step a1 b1,
// a1 b2 with A from the operand collector, a2 b1).  8 epilogue warps read TMEM (the last column batches are parked in
// shared memory first so that TMEM is released early), apply the closed form with the FP64-accumulated window sums on
// the packed FP32 pipe, blend, map to a colour and store the pixel and its mirror (the mirror goes through a small
// shared-memory transpose so that both stores are coalesced).  DESIGN.md section 4 (K4) has the measurements.
#pragma once
#include <cuda_fp16.h>

#include "corr_tc.cuh"
#include "selfsim_fast.cuh"

namespace sgz {

constexpr int kSgBuildWarps = 7, kSgEpiWarps = 8;   // warp 0 = record producer, warps 1..6 = builders (expansion mode only)
constexpr int kSgBuilders = kSgBuildWarps - 1;
constexpr int kSgThreads = (kSgBuildWarps + 1 + kSgEpiWarps) * 32;   // 512: the register file is handed out per 4 warps
constexpr int kSgPPitch = 17;                                        // transpose buffer: 32 rows x 16 pixels, padded

struct SelfTcGeom {
  int H, nks, nSlab, slabKs;  // K steps of 16; a ring stage ("slab") holds slabKs of them
  int g, dp, kcStep, span;    // record grid (see above); span = records one operand stage holds
  int nStage;                 // ring depth (slabs)
  int nRecStage;              // record stages (channels requested ahead + 1)
  int dump;                   // accumulator batches (32 KB each) the epilogue may park in shared memory
  int aDesc;                  // decim | 8: A and B are read in place from the record stages
  uint32_t matBytes, stageBytes;
  uint32_t recPartBytes, recStageBytes, tailBytes;
  size_t smemBytes;
  bool ok;
};

inline SelfTcGeom self_tc_geom(int H, int decim, size_t smemLimit, bool allowADesc) {
  SelfTcGeom G{};
  G.H = H;
  G.nks = (H + 15) / 16;
  G.g = decim & -decim;
  if (G.g > 8) G.g = 8;
  G.dp = decim / G.g;
  G.kcStep = 8 / G.g;
  G.span = G.dp * 127 + G.kcStep * (2 * G.nks - 1) + 1;
  G.recPartBytes = (uint32_t)G.span * 16u;
  G.recStageBytes = (4u * G.recPartBytes + 127u) / 128u * 128u;
  G.aDesc = allowADesc && G.dp == 1;
  G.ok = false;
  G.tailBytes = 0;
  if (G.aDesc) {
    // decim | 8: both operands are read in place from the record stages.  When H is not a multiple of 16 the last K step
    // of B must end at the window's end: its masked copy comes pre-built from global memory (k_self_records) as part of
    // the record stage, in canonical [chunk][row] order.  No builder warps, no ring.
    G.tailBytes = (H & 15) ? 4u * 2048u : 0u;                      // B first / second part x 2 chunks x 128 rows x 16 B
    G.recStageBytes = (4u * G.recPartBytes + G.tailBytes + 127u) / 128u * 128u;
    G.slabKs = G.nks; G.nSlab = 1; G.nStage = 0; G.matBytes = 0; G.stageBytes = 0;
    const size_t fixed = (size_t)kSgEpiWarps * 32 * kSgPPitch * 4 + 128 * 16 + 512 + 1024;
    // shared memory = record ring + parked accumulators: as many parked batches as leave a ring of >= 5 stages
    G.dump = 0;
    for (int d = 3; d >= 0; d--) {
      const size_t left = smemLimit - fixed - (size_t)d * 32768;
      if (smemLimit >= fixed + (size_t)d * 32768 && left / G.recStageBytes >= (size_t)(d > 0 ? 5 : 3)) { G.dump = d; break; }
    }
    G.nRecStage = (int)std::min<size_t>(8, (smemLimit - fixed - (size_t)G.dump * 32768) / G.recStageBytes);
    G.smemBytes = fixed + (size_t)G.dump * 32768 + (size_t)G.nRecStage * G.recStageBytes;
    G.ok = G.nRecStage >= 3;
    if (G.ok) return G;
    G.aDesc = 0;
    G.tailBytes = 0;
    G.recStageBytes = (4u * G.recPartBytes + 127u) / 128u * 128u;
  }
  // expansion mode: large ring stages mean few barrier round trips and fences (at least 3 stages if they fit, else smaller
  // slabs); two record stages are enough for the builders, more only if shared memory is left over
  {
    const size_t fixed = (size_t)kSgEpiWarps * 32 * kSgPPitch * 4 + 128 * 16 + 512 + 1024;
    for (int ks = std::min(G.nks, 4); ks >= 1 && !G.ok; ks--) {
      const uint32_t matBytes = (uint32_t)ks * 4096u, stageBytes = 4u * matBytes;
      for (int ns = 4; ns >= (ks > 2 ? 3 : 2); ns--) {
        if (fixed + 2 * (size_t)G.recStageBytes + (size_t)ns * stageBytes <= smemLimit) {
          G.slabKs = ks; G.nSlab = (G.nks + ks - 1) / ks; G.nStage = ns; G.nRecStage = 2;
          G.matBytes = matBytes; G.stageBytes = stageBytes;
          while (G.nRecStage < 4 && fixed + (size_t)(G.nRecStage + 1) * G.recStageBytes + (size_t)ns * stageBytes <= smemLimit)
            G.nRecStage++;
          G.smemBytes = fixed + (size_t)G.nRecStage * G.recStageBytes + (size_t)ns * stageBytes;
          G.ok = true;
          break;
        }
      }
    }
  }
  return G;
}

struct SelfTcParams {
  SelfFastParams f;          // geometry, colours, rgb, shifts; ws1 / ws2 are NOT used (the scaled sums below are)
  const uint4 *rec1, *rec2;  // records [part][channel][nRec] of file 1 / file 2
  int64_t nRec;
  const float4 *wsA, *wsB;   // per window (S_T, Q_T, S_S, Q_S) of the scaled centred data (file 1 / file 2)
  const int2 *tiles;
  int nTiles;
  int nks, nSlab, slabKs, dp, kcStep, span, nStage, nRecStage;
  uint32_t matBytes, stageBytes, recPartBytes, recStageBytes, tailBytes;
  const uint4 *tail2;        // pre-masked last K step of file 2's windows: [part][channel][chunk 0/1][nTailRows] (in-place mode)
  int64_t nTailRows;
  float *simMat;             // optional [imgExt][imgExt] raw sims at (a, b >= a) (parity checks), or nullptr
  int aDesc;                 // decim | 8: the A operand is read straight from the record stage (see the issuer)
  long long *prof;           // developer probe (SGZ_SELF_TC_PROF): per CTA 16 cycle counters, or nullptr
  // Passes.  With both groups in play (0 < temporalWeight < 1) the image takes two launches so that the accumulators of
  // a launch fit TMEM twice and the epilogue of tile n overlaps the MMAs of tile n + 1:
  //   pass 1 (doT, storeT): temporal group only, one 128-column region per tile (4 in flight), writes the temporal
  //                         coefficient of every cell to corrT[c][a];
  //   pass 2 (doS, loadT):  spectral group, main + correction regions (2 tiles in flight), reads corrT, writes pixels.
  // A single group (weight 0 or 1) is one launch with doT or doS alone.
  int storeT, loadT;         // (which groups a launch computes is the kernel's template argument)
  int dump;                  // batches of accumulators the epilogue parks in shared memory to release TMEM early (0..3)
  int twoMain;               // long windows: the spectral main products alternate between two regions by channel parity,
                             // halving the chain (and the truncation bias) per accumulator; no spare region then
  float *corrT;              // [imgExt][imgExt], element [c][a] (column-major so that lanes = rows a are coalesced)
  // Long windows (chains beyond the error budget of one accumulator): the window is cut into chunks of H frames, one
  // launch per chunk.  A launch reads the records recOff further on (its chunk starts recOff * g frames into the window),
  // adds the raw Gram sums of the earlier chunks (loadG) and either parks its own sums (storeG) or -- last chunk --
  // applies the closed form with the window sums and the length Hfull of the WHOLE window.
  int64_t recOff;
  int Hfull;                 // 0: f.base.H is the whole window
  float *gPart;              // [2][imgExt][imgExt] temporal / spectral partial Gram sums, element [c][a]
  int storeG, loadG;
};

// max |x| per group (as float bits; |NaN| compares above everything) next to the means of k_self_means
__global__ void k_self_absmax(const float *__restrict__ x, int64_t stride, int64_t n, int numCh, unsigned int *out) {
  unsigned int mT = 0, mS = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    mT = max(mT, __float_as_uint(x[i]) & 0x7fffffffu);
    for (int c = 1; c < numCh; c++) mS = max(mS, __float_as_uint(x[(int64_t)c * stride + i]) & 0x7fffffffu);
  }
  for (int d = 16; d > 0; d >>= 1) {
    mT = max(mT, __shfl_xor_sync(0xffffffffu, mT, d));
    mS = max(mS, __shfl_xor_sync(0xffffffffu, mS, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(out, mT); atomicMax(out + 1, mS); }
}

// records: rec[part][c][rho] = FP16 first / second part of scale * (x[c][g rho + j] - shift_group), j = 0..7; zero past nValid
__global__ void k_self_records(const float *__restrict__ x, int64_t stride, int64_t nValid, int numCh, int g, int64_t nRec,
                               float shiftT, float shiftS, float scale, uint4 *__restrict__ rec) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)numCh * nRec) return;
  const int c = (int)(idx / nRec);
  const int64_t rho = idx - (int64_t)c * nRec;
  const float sh = c == 0 ? shiftT : shiftS;
  __half hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int64_t fr = (int64_t)g * rho + j;
    const float v = fr < nValid ? __fmul_rn(__fsub_rn(x[(int64_t)c * stride + fr], sh), scale) : 0.f;
    hi[j] = __float2half_rn(v);
    lo[j] = __float2half_rn(v - __half2float(hi[j]));
  }
  rec[idx] = *reinterpret_cast<const uint4 *>(hi);
  rec[(int64_t)numCh * nRec + idx] = *reinterpret_cast<const uint4 *>(lo);
}

// In-place mode, H % 16 != 0: the last K step (chunks kc0 = 2 (nks - 1) and kc0 + 1) of every window row of file 2, cut
// off at the window's end: tail[part][c][kcl][row] = record (row + kcStep (kc0 + kcl)) with the halves k >= H zeroed
__global__ void k_self_tail(const uint4 *__restrict__ rec, int numCh, int64_t nRec, int kcStep, int kc0, int H, int64_t nRows,
                            uint4 *__restrict__ tail, int64_t recOff = 0) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= 2 * (int64_t)numCh * 2 * nRows) return;
  const int64_t row = idx % nRows;
  const int kcl = (int)((idx / nRows) & 1);
  const int64_t pc = idx / (2 * nRows);               // part * numCh + c
  const int64_t rho = row + (int64_t)kcStep * (kc0 + kcl) + recOff;   // (chunked windows start recOff records in)
  uint4 v = rho < nRec ? rec[pc * nRec + rho] : make_uint4(0, 0, 0, 0);
  const int nv = H - 8 * (kc0 + kcl);
  if (nv < 8) {
    v.x = nv >= 2 ? v.x : (nv == 1 ? v.x & 0xffffu : 0u);
    v.y = nv >= 4 ? v.y : (nv == 3 ? v.y & 0xffffu : 0u);
    v.z = nv >= 6 ? v.z : (nv == 5 ? v.z & 0xffffu : 0u);
    v.w = nv == 7 ? v.w & 0xffffu : 0u;
  }
  tail[idx] = v;
}

// (S, Q) of every decimated window of the scaled centred data, FP64 accumulation, both groups in one float4
__global__ void k_self_wsums4(const float *__restrict__ x, int64_t stride, int numCh, int H, int decim, int imgExt,
                              float shiftT, float shiftS, float scale, float4 *__restrict__ ws) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= imgExt) return;
  const int64_t f0 = (int64_t)decim * a;
  double s = 0, q = 0;
  for (int h = 0; h < H; h++) { const double v = (double)__fmul_rn(__fsub_rn(x[f0 + h], shiftT), scale); s += v; q += v * v; }
  float4 o;
  o.x = (float)s; o.y = (float)q;
  s = 0; q = 0;
  for (int c = 1; c < numCh; c++)
    for (int h = 0; h < H; h++) {
      const double v = (double)__fmul_rn(__fsub_rn(x[(int64_t)c * stride + f0 + h], shiftS), scale);
      s += v; q += v * v;
    }
  o.z = (float)s; o.w = (float)q;
  ws[a] = o;
}

// no-swizzle K-major operand: rows of a core matrix 16 B apart, 8-row groups SBO apart, the two 8-element K chunks of
// one K step LBO apart
__device__ __forceinline__ uint64_t sg_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma_acc(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc) {   // D += A * B
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmemD),
      "l"(da), "l"(db), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void sg_epi_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

__device__ __forceinline__ float sg_coeff(float G, float Sa, float Qa, float Sb, float Qb, float inv4N) {
  const float S = Sa + Sb;
  const float T = S * S * inv4N;
  const float den = fmaf(0.5f, Qa + Qb, -T);
  return __fdividef(G - T, den);   // 0/0 -> NaN like the reference (constant windows)
}

// the same with halved sums of squares: hQ = Q / 2
__device__ __forceinline__ float sg_coeff_h(float G, float Sa, float hQa, float Sb, float hQb, float inv4N) {
  const float S = Sa + Sb;
  const float t1 = S * inv4N;
  const float den = fmaf(-t1, S, hQa + hQb);
  float r;                                    // MUFU.RCP alone (__fdividef adds range scaling the operands never need):
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));   // den = 0 -> inf, and 0 * inf = NaN like the reference's 0/0
  return fmaf(-t1, S, G) * r;
}

// both groups at once on the packed FP32 pipe (FADD2 / FMUL2 / FFMA2): x = temporal, y = spectral
__device__ __forceinline__ float2 sg_coeff2(float2 G, float2 Sa, float2 hQa, float2 Sb, float2 hQb, float2 negInv4N) {
  const float2 S = __fadd2_rn(Sa, Sb);
  const float2 nt1 = __fmul2_rn(S, negInv4N);                 // -S / 4N
  const float2 den = __ffma2_rn(nt1, S, __fadd2_rn(hQa, hQb));
  const float2 num = __ffma2_rn(nt1, S, G);
  float rx, ry;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(den.y));
  return __fmul2_rn(num, make_float2(rx, ry));
}

// GrayScale with colorWarp = 1 without FP64: clamp(floor(s * 255 + 0.5), 0, 255) computed exactly -- s * 255 = k + fr with
// k = trunc, fr = exact remainder, so the reference's (int)((double) f + 0.5) is k + (fr >= 0.5); NaN -> 0 like d2i_java
__device__ __forceinline__ int32_t sg_grey(float sim, float colorScale, int colorInv) {
  const float v = __fmul_rn(fmaxf(sim, 0.f), colorScale);
  const float s = colorInv ? __fsub_rn(1.0f, v) : v;
  const float f = fminf(fmaxf(__fmul_rn(s, 255.0f), 0.f), 256.f);
  int k = __float2int_rz(f);
  k += (f - (float)k) >= 0.5f ? 1 : 0;
  k = (sim != sim) ? 0 : min(k, 255);
  return k * 0x010101;
}

// colour of one cell; identical to self_color() except that pow(m, 1.0) is skipped (it returns m exactly)
__device__ __forceinline__ int32_t sg_color(const SelfParams &p, float sim, bool warpOne) {
  if (!warpOne || p.lut != nullptr) return self_color(p, sim);
  float m = (sim != sim) ? sim : fmaxf(0.0f, sim);
  if (sim == 0.0f) m = 0.0f;
  const float v = __fmul_rn(m, p.colorScale);
  const float s = p.colorInv ? __fsub_rn(1.0f, v) : v;
  int32_t i = d2i_java(__dadd_rn((double)__fmul_rn(s, 255.0f), 0.5));
  i = max(0, min(255, i));
  return i * 0x010101;
}

// kMode: 0 = both groups in one launch, 1 = temporal group only, 2 = spectral group only
template <int kMode>
__global__ void __launch_bounds__(kSgThreads, 1) k_self_gram_tc(const SelfTcParams p) {
  constexpr bool kDoT = kMode != 2, kDoS = kMode != 1;
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  const SelfParams &b = p.f.base;
  unsigned char *base = smemRaw + ((1024 - (smem_u32(smemRaw) & 1023)) & 1023);
  unsigned char *ring = base;                                               // [nStage][mats][2 slabKs chunks][128 rows][16 B]
  unsigned char *recBase = ring + (size_t)p.nStage * p.stageBytes;          // [nRecStage][A1, A2, B1, B2][span] records
  int32_t *P = reinterpret_cast<int32_t *>(recBase + (size_t)p.nRecStage * p.recStageBytes);   // [8 warps][32][17]
  float4 *colW = reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(P) + (size_t)kSgEpiWarps * 32 * kSgPPitch * 4);
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(colW) + 128 * 16);
  uint64_t *full = bars, *empty = bars + 4, *recFull = bars + 8, *recEmpty = bars + 16, *accFull = bars + 24,
           *accEmpty = bars + 28;
  uint32_t *tmemSlot = reinterpret_cast<uint32_t *>(bars + 32);
  float *dump = reinterpret_cast<float *>(bars + 64);                        // [dump][16 columns][T, S][256 epilogue threads]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 4; s++) {
      mbar_init(full + s, kSgBuilders); mbar_init(empty + s, 1);
      mbar_init(accFull + s, 1); mbar_init(accEmpty + s, kSgEpiWarps);
    }
    // a record stage is released by the MMAs that read it (in-place mode) or by the builders that expanded it
    for (int s = 0; s < 8; s++) { mbar_init(recFull + s, 1); mbar_init(recEmpty + s, p.aDesc ? 1 : kSgBuilders); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmemSlot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmemSlot;

  // channels of this pass and its TMEM staging: temporal = one region per tile, spectral = main + correction regions
  const int cBegin = kDoT ? 0 : 1, cEnd = kDoS ? b.numCh : 1;
  const int accStages = kDoS ? ((kDoT || p.twoMain) ? 1 : 2) : 4;
  // TMEM columns inside a stage.  Temporal region: column 0 (alternating with the spare region 384 in the one-launch mode,
  // see the issuer); spectral main products: regMa (and regMb for even channels with twoMain); corrections: regC.
  const uint32_t regMa = kDoT ? 128u : 0u, regMb = regMa + 128u, regC = regMa + (p.twoMain ? 256u : 128u);
  const bool spareT = kMode == 0 && !p.twoMain;
  const uint32_t accCols = 512u / (uint32_t)accStages;
  const int H = b.H;
  const uint32_t matUnits = p.matBytes / 16;

  if (warp == 0) {
    // =========================== record producer ===========================
    // the records both operands of the n-th (tile, channel) of this CTA need (+ B's pre-masked last K step in in-place
    // mode); runs as far ahead as the record stages allow
    const int64_t chanRecs = p.nRec, partRecs = (int64_t)b.numCh * p.nRec;
    uint32_t n = 0;
    for (int t = blockIdx.x; t < p.nTiles; t += gridDim.x) {
      const int2 tl = p.tiles[t];
      for (int c = cBegin; c < cEnd; c++, n++) {
        const int rs = n % p.nRecStage;
        const uint32_t use = n / p.nRecStage;
        if (use > 0) tc_wait(recEmpty + rs, (use - 1) & 1);
        if (tc_elect()) {
          unsigned char *dst = recBase + (size_t)rs * p.recStageBytes;
          mbar_expect_tx(recFull + rs, 4u * p.recPartBytes + p.tailBytes);
          const uint4 *a0 = p.rec1 + (int64_t)c * chanRecs + (int64_t)p.dp * tl.x + p.recOff;
          const uint4 *b0 = p.rec2 + (int64_t)c * chanRecs + (int64_t)p.dp * tl.y + p.recOff;
          bulk_g2s(dst, a0, p.recPartBytes, recFull + rs);
          bulk_g2s(dst + p.recPartBytes, a0 + partRecs, p.recPartBytes, recFull + rs);
          bulk_g2s(dst + 2 * p.recPartBytes, b0, p.recPartBytes, recFull + rs);
          bulk_g2s(dst + 3 * p.recPartBytes, b0 + partRecs, p.recPartBytes, recFull + rs);
          if (p.tailBytes) {   // [part][chunk][128 rows]
            unsigned char *td = dst + 4 * p.recPartBytes;
            for (int part = 0; part < 2; part++)
              for (int kcl = 0; kcl < 2; kcl++)
                bulk_g2s(td + (part * 2 + kcl) * 2048,
                         p.tail2 + (((int64_t)part * b.numCh + c) * 2 + kcl) * p.nTailRows + tl.y, 2048, recFull + rs);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < kSgBuildWarps) {
    // =========================== builders (expansion mode: decim does not divide 8) ===========================
    const int bw = warp - 1;
    uint32_t slabCtr = 0, chCtr = 0;
    long long bRec = 0, bEmpty = 0, bBuild = 0, bTot = clock64(), tB;
    for (int t = blockIdx.x; t < p.nTiles && !p.aDesc; t += gridDim.x) {
      for (int c = cBegin; c < cEnd; c++, chCtr++) {
        const int rs = chCtr % p.nRecStage;
        tB = clock64();
        tc_wait(recFull + rs, (chCtr / p.nRecStage) & 1);
        bRec += clock64() - tB;
        const uint4 *recS = reinterpret_cast<const uint4 *>(recBase + (size_t)rs * p.recStageBytes);
        for (int slab = 0; slab < p.nSlab; slab++, slabCtr++) {
          const int st = slabCtr % p.nStage;
          const uint32_t use = slabCtr / p.nStage;
          tB = clock64();
          if (use > 0) tc_wait(empty + st, (use - 1) & 1);   // the MMAs that read this stage are done
          bEmpty += clock64() - tB; tB = clock64();
          uint4 *dstS = reinterpret_cast<uint4 *>(ring + (size_t)st * p.stageBytes);
          const int kc0 = 2 * p.slabKs * slab;
          const int nkcl = 2 * min(p.slabKs, p.nks - p.slabKs * slab);     // 8-element chunks of this stage
          const int nPairs = 4 * nkcl;                                    // (operand part, chunk) pairs x 4 row groups
          // one pair per step and warp: the four row groups are four independent 16-byte copies per lane
          for (int pi = bw; pi < nPairs; pi += kSgBuilders) {
            int mi = 0, kcl = pi;
            while (kcl >= nkcl) { kcl -= nkcl; mi++; }
            const int kc = kc0 + kcl;
            const uint4 *src = recS + (size_t)mi * p.span + (size_t)p.kcStep * kc + (size_t)p.dp * lane;
            uint4 *dst = dstS + (size_t)mi * matUnits + (size_t)kcl * 128 + lane;
            const int step = 32 * p.dp;
            uint4 v0 = src[0], v1 = src[step], v2 = src[2 * step], v3 = src[3 * step];
            const int nv = H - 8 * kc;                       // valid k of this chunk; beyond H the window has ended
            if (nv < 8) {
              const uint32_t mx = nv >= 2 ? 0xffffffffu : (nv == 1 ? 0xffffu : 0u), my = nv >= 4 ? 0xffffffffu : (nv == 3 ? 0xffffu : 0u);
              const uint32_t mz = nv >= 6 ? 0xffffffffu : (nv == 5 ? 0xffffu : 0u), mw = nv == 7 ? 0xffffu : 0u;
              v0.x &= mx; v0.y &= my; v0.z &= mz; v0.w &= mw;
              v1.x &= mx; v1.y &= my; v1.z &= mz; v1.w &= mw;
              v2.x &= mx; v2.y &= my; v2.z &= mz; v2.w &= mw;
              v3.x &= mx; v3.y &= my; v3.z &= mz; v3.w &= mw;
            }
            dst[0] = v0; dst[32] = v1; dst[64] = v2; dst[96] = v3;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
          __syncwarp();
          if (lane == 0) mbar_arrive(full + st);
          bBuild += clock64() - tB;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(recEmpty + rs);
      }
    }
    if (p.prof && warp == 1 && lane == 0) {
      long long *o = p.prof + 16 * blockIdx.x + 8;
      o[0] = clock64() - bTot; o[1] = bRec; o[2] = bEmpty; o[3] = bBuild;
    }
  } else if (warp == kSgBuildWarps) {
    // =========================== record producer + MMA issuer ===========================
    // D = F32, A = B = F16, both K-major, M = 128, N = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t slabCtr = 0, chCtr = 0, tileIt = 0;
    const bool prof = p.prof != nullptr;
    long long iRec = 0, iAcc = 0, iFull = 0, iIssue = 0, iTot = clock64(), tI = 0;
    if (p.aDesc) {
      // ---- in-place mode.  decim | 8: a window starts every record (dp = 1), so the 8 rows of a core matrix ARE 8
      // consecutive records (16 B apart), row groups are 128 B apart and chunk kc of a row lies kcStep records further:
      // the record stage itself is the K-major operand of BOTH sides, with overlapping core matrices (LBO = kcStep * 16
      // B).  Only B's last K step is a separate, pre-masked block of the stage when H is not a multiple of 16.
      // The queue of issued MMAs is short, so the loop keeps the gap between two channels small: descriptors are stage 0's
      // plus a stage offset, waits come first, nothing else sits between the MMAs of consecutive channels.
      const uint32_t lbo = (uint32_t)p.kcStep * 16u, r0 = smem_u32(recBase);
      const uint64_t partU = p.recPartBytes >> 4, stageU = p.recStageBytes >> 4, inc = 2u * (uint32_t)p.kcStep;
      const uint64_t A1 = sg_desc(r0, lbo, 128), T1 = sg_desc(r0 + 4 * p.recPartBytes, 2048, 128);
      const int nks = p.nks, nIn = (H & 15) ? nks - 1 : nks;     // K steps whose B is read in place
      uint32_t rs = 0, rsPar = 0;
      for (int t = blockIdx.x; t < p.nTiles; t += gridDim.x, tileIt++) {
        const int as = tileIt % accStages;
        const uint32_t accBase = tmem + (uint32_t)as * accCols;
        for (int c = cBegin; c < cEnd; c++) {
          if (prof) tI = clock64();
          tc_wait<false>(recFull + rs, rsPar);
          if (prof) { iRec += clock64() - tI; tI = clock64(); }
          // One-launch mode with both groups: the temporal region alternates between columns 0 and 384, so the 18 temporal
          // MMAs of tile n + 1 run while the epilogue still reads tile n (the region they overwrite was drained with tile
          // n - 1); only the spectral regions wait for the epilogue.
          const bool waitHere = spareT ? c == 1 : c == cBegin;
          if (waitHere && tileIt >= (uint32_t)accStages)
            tc_wait<false>(accEmpty + as, ((tileIt / accStages) - 1) & 1);   // the epilogue has drained this TMEM stage
          if (prof) { iAcc += clock64() - tI; tI = clock64(); }
          asm volatile("tcgen05.fence::after_thread_sync;");
          // regions of the stage: temporal = one region for all three products; spectral = main + correction, after the
          // temporal region when both groups share a launch
          const uint32_t dMain = accBase + (c == 0 ? ((spareT && (tileIt & 1)) ? 384u : 0u)
                                                   : ((p.twoMain && (c & 1) == 0) ? regMb : regMa));
          const uint32_t dCorr = c == 0 ? dMain : accBase + regC;
          const uint64_t a1 = A1 + stageU * rs, a2 = a1 + partU, b1 = a1 + 2 * partU, b2 = a1 + 3 * partU;
          const uint64_t t1 = T1 + stageU * rs, t2 = t1 + 256;
          // first MMA into a region of this tile overwrites
          const uint32_t accFirst = (c == 0 || c == 1 || (p.twoMain && c == 2)) ? 0u : 1u, accFirstC = c == 1 ? 0u : 1u;
          if (tc_elect()) {
            // per K step: a1 b1 -> main, a1 b2 -> correction (A taken from the collector, not from shared memory again),
            // a2 b1 -> correction.  (Changing the accumulator between MMAs costs nothing: tools/umma_rate_probe.cu.)
            for (int ks = 0; ks < nks; ks++) {
              const uint64_t x1 = a1 + inc * ks, x2 = a2 + inc * ks;
              const uint64_t y1 = ks < nIn ? b1 + inc * ks : t1, y2 = ks < nIn ? b2 + inc * ks : t2;
              tc_mma_fill(dMain, x1, y1, idesc, ks > 0 ? 1u : accFirst);
              tc_mma_lastuse(dCorr, x1, y2, idesc, (ks > 0 || c == 0) ? 1u : accFirstC);
              tc_mma_acc(dCorr, x2, y1, idesc);
            }
            tc_commit(recEmpty + rs);                              // the MMAs read the record stage
            if (c + 1 == cEnd) tc_commit(accFull + as);
          }
          __syncwarp();
          if (++rs == (uint32_t)p.nRecStage) { rs = 0; rsPar ^= 1; }
          if (prof) iIssue += clock64() - tI;
        }
      }
    }
    for (int t = blockIdx.x; t < p.nTiles && !p.aDesc; t += gridDim.x, tileIt++) {
      // ---- expansion mode: operands come from the ring the builders fill
      const int as = tileIt % accStages;
      const uint32_t accBase = tmem + (uint32_t)as * accCols;
      for (int c = cBegin; c < cEnd; c++, chCtr++) {
        tI = clock64();
        if (c == cBegin && tileIt >= (uint32_t)accStages)
          tc_wait<false>(accEmpty + as, ((tileIt / accStages) - 1) & 1);   // the epilogue has drained this TMEM stage
        iAcc += clock64() - tI;
        const uint32_t dMain = accBase + (c == 0 ? ((spareT && (tileIt & 1)) ? 384u : 0u)
                                                 : ((p.twoMain && (c & 1) == 0) ? regMb : regMa));
        const uint32_t dCorr = c == 0 ? dMain : accBase + regC;
        for (int slab = 0; slab < p.nSlab; slab++, slabCtr++) {
          const int st = slabCtr % p.nStage;
          const uint32_t use = slabCtr / p.nStage;
          tI = clock64();
          tc_wait<false>(full + st, use & 1);
          iFull += clock64() - tI; tI = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;");
          if (tc_elect()) {
            const uint32_t sb = smem_u32(ring + (size_t)st * p.stageBytes);
            // first MMA into a region of this tile overwrites; the temporal group keeps all three products in one region
            const bool first = slab == 0 && (c == 0 || c == 1 || (p.twoMain && c == 2)), firstC = slab == 0 && c == 1;
            // descriptors of the first K step; a K step further is a constant increment of the start-address fields
            const int nksHere = min(p.slabKs, p.nks - p.slabKs * slab);
            const uint64_t a1 = sg_desc(sb, 2048, 128), a2 = sg_desc(sb + p.matBytes, 2048, 128);
            const uint64_t b1 = sg_desc(sb + 2 * p.matBytes, 2048, 128), b2 = sg_desc(sb + 3 * p.matBytes, 2048, 128);
            for (int ks = 0; ks < nksHere; ks++)             // 2 chunks x 2048 B = 256 units per K step
              tc_mma(dMain, a1 + 256u * ks, b1 + 256u * ks, idesc, !(first && ks == 0));
            for (int ks = 0; ks < nksHere; ks++) {
              tc_mma(dCorr, a2 + 256u * ks, b1 + 256u * ks, idesc, c == 0 ? 1u : !(firstC && ks == 0));
              tc_mma_acc(dCorr, a1 + 256u * ks, b2 + 256u * ks, idesc);
            }
            tc_commit(empty + st);
            if (c + 1 == cEnd && slab + 1 == p.nSlab) tc_commit(accFull + as);
          }
          __syncwarp();
          iIssue += clock64() - tI;
        }
      }
    }
    if (p.prof && lane == 0) {
      long long *o = p.prof + 16 * blockIdx.x;
      o[0] = clock64() - iTot; o[1] = iRec; o[2] = iAcc; o[3] = iFull; o[4] = iIssue; o[5] = tileIt;
    }
  } else {
    // =========================== epilogue ===========================
    // warp -> TMEM lane quarter (rows of the tile) and column half; thread = one row a, 64 columns in batches of 16
    const int ew = warp - (kSgBuildWarps + 1), quarter = warp & 3, half = ew >> 2, et = ew * 32 + lane;
    int32_t *Pw = P + (size_t)ew * 32 * kSgPPitch;
    const int ext = b.imgExt;
    const int Hf = p.Hfull > 0 ? p.Hfull : H;      // the closed form is that of the whole window
    const float invNT = (float)(1.0 / (4.0 * (double)Hf)), invNS = (float)(1.0 / (4.0 * (double)(b.numCh - 1) * (double)Hf));
    const bool warpOne = b.colorWarp == 1.0f;
    const float wT = b.weight, wS = __fsub_rn(1.0f, b.weight);
    const bool useT = kDoT, useS = kDoS;
    // TMEM columns of the regions inside a stage: one-launch two-group mode [T | S main | S corr], else [main | corr]
    const uint32_t colT = 0u, colM = regMa, colC = regC;
    uint32_t tileIt = 0;
    long long eAcc = 0, eMain = 0, tE;
    for (int t = blockIdx.x; t < p.nTiles; t += gridDim.x, tileIt++) {
      const int as = tileIt % accStages;
      const int2 tl = p.tiles[t];
      const int ta = tl.x, tb = tl.y;
      sg_epi_sync();                                       // everybody is done with the previous tile's column sums
      if (et < 128) {                                      // column sums as (S_T, S_S, Q_T / 2, Q_S / 2)
        const float4 w = tb + et < ext ? p.wsB[tb + et] : make_float4(0.f, 0.f, 0.f, 0.f);
        colW[et] = make_float4(w.x, w.z, 0.5f * w.y, 0.5f * w.w);
      }
      sg_epi_sync();
      const int a = ta + quarter * 32 + lane;
      float4 wa;
      {
        const float4 w = a < ext ? p.wsA[a] : make_float4(0.f, 0.f, 0.f, 0.f);
        wa = make_float4(w.x, w.z, 0.5f * w.y, 0.5f * w.w);
      }
      const float2 waS = make_float2(wa.x, wa.y), waQ = make_float2(wa.z, wa.w);
      const float2 negInv = make_float2(-invNT, -invNS);
      const bool rowOk = a < ext && a >= b.colBegin && a < b.colEnd;
      // interior tile (all 128 x 128 cells exist, strictly above the diagonal): no per-cell predicates, 32-bit pixel
      // offsets (imgExt <= 0xB504, so imgExt^2 < 2^31); the pixel path also needs the plain grey scale
      const bool interior = ta + 128 <= ext && tb + 128 <= ext && tb >= ta + 128 && ta >= b.colBegin && ta + 128 <= b.colEnd;
      const bool fast = interior && warpOne && b.lut == nullptr && p.simMat == nullptr;
      tE = clock64();
      tc_wait(accFull + as, (tileIt / accStages) & 1);
      eAcc += clock64() - tE; tE = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t laneAddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)as * accCols + 64u * (uint32_t)half;
      // TMEM holds ONE tile in the one-launch mode, so the next tile's MMAs wait until the last accumulator has been read.
      // To shorten that, the accumulators of the last `dump` batches are first copied to shared memory (a thread reads
      // back only what it wrote: no synchronisation), the TMEM stage is released after the load of batch 3 - dump, and
      // the dumped batches are evaluated from shared memory while the MMAs of the next tile already run.
      const int D = p.dump;
      const uint32_t colTt = (spareT && (tileIt & 1)) ? 384u : colT;
#pragma unroll 1
      for (int bt = 4 - D; bt < 4; bt++) {
        uint32_t uT[16], uM[16], uC[16];
        if (useT) tc_ld16_nowait(laneAddr + colTt + 16u * bt, uT);
        if (useS) {
          tc_ld16_nowait(laneAddr + colM + 16u * bt, uM);
          tc_ld16_nowait(laneAddr + colC + 16u * bt, uC);
        }
        tc_ld_wait();
        if (useS && p.twoMain) {   // second main region (even channels)
#pragma unroll
          for (int i = 0; i < 16; i++) uM[i] = __float_as_uint(__uint_as_float(uM[i]) + __uint_as_float(uC[i]));
          tc_ld16_nowait(laneAddr + regMb + 16u * bt, uC);
          tc_ld_wait();
        }
        float *dp = dump + (size_t)(bt - (4 - D)) * (2 * 16 * kSgEpiWarps * 32) + et;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (useT) dp[(2 * i) * (kSgEpiWarps * 32)] = __uint_as_float(uT[i]);
          if (useS) dp[(2 * i + 1) * (kSgEpiWarps * 32)] = __uint_as_float(uM[i]) + __uint_as_float(uC[i]);
        }
      }
#pragma unroll 1
      for (int bt = 0; bt < 4; bt++) {
        uint32_t uT[16], uM[16], uC[16];
        if (bt < 4 - D) {
          if (useT) tc_ld16_nowait(laneAddr + colTt + 16u * bt, uT);
          if (useS) {
            tc_ld16_nowait(laneAddr + colM + 16u * bt, uM);
            tc_ld16_nowait(laneAddr + colC + 16u * bt, uC);
          }
          tc_ld_wait();
          if (useS && p.twoMain) {   // second main region (even channels)
#pragma unroll
            for (int i = 0; i < 16; i++) uM[i] = __float_as_uint(__uint_as_float(uM[i]) + __uint_as_float(uC[i]));
            tc_ld16_nowait(laneAddr + regMb + 16u * bt, uC);
            tc_ld_wait();
          }
          if (useS) {      // spectral Gram = main + correction products
#pragma unroll
            for (int i = 0; i < 16; i++) uM[i] = __float_as_uint(__uint_as_float(uM[i]) + __uint_as_float(uC[i]));
          }
          if (bt == 3 - D) {   // the last accumulators are out of TMEM: hand the stage back to the issuer
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty + as);
          }
        } else {
          const float *dp = dump + (size_t)(bt - (4 - D)) * (2 * 16 * kSgEpiWarps * 32) + et;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            if (useT) uT[i] = __float_as_uint(dp[(2 * i) * (kSgEpiWarps * 32)]);
            if (useS) uM[i] = __float_as_uint(dp[(2 * i + 1) * (kSgEpiWarps * 32)]);
          }
        }
        const int c0 = tb + 64 * half + 16 * bt;
        const float4 *cw = colW + 64 * half + 16 * bt;
        if (p.loadG || p.storeG) {    // chunked window: Gram sums of the chunks so far
          float *gT = p.gPart, *gS = p.gPart + (int64_t)ext * ext;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const bool ok = interior || (a < ext && c0 + i < ext);
            const int64_t o = (int64_t)(c0 + i) * ext + a;
            if (p.loadG && ok) {
              if (useT) uT[i] = __float_as_uint(__uint_as_float(uT[i]) + __ldcs(gT + o));
              if (useS) uM[i] = __float_as_uint(__uint_as_float(uM[i]) + __ldcs(gS + o));
            }
            if (p.storeG && ok) {
              if (useT) __stcs(gT + o, __uint_as_float(uT[i]));
              if (useS) __stcs(gS + o, __uint_as_float(uM[i]));
            }
          }
          if (p.storeG) continue;
        }
        if (kMode == 1 && p.storeT) {
          // pass 1 of 2: the temporal coefficient of every cell of the tile -> corrT[c][a] (lanes = rows a: coalesced)
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float4 wb = cw[i];
            const float temporal = sg_coeff_h(__uint_as_float(uT[i]), wa.x, wa.z, wb.x, wb.z, invNT);
            if (interior || (a < ext && c0 + i < ext)) p.corrT[(int64_t)(c0 + i) * ext + a] = temporal;
          }
          continue;
        }
        float tv[16];
        if (kMode == 2 && p.loadT) {
#pragma unroll
          for (int i = 0; i < 16; i++) tv[i] = (interior || (a < ext && c0 + i < ext)) ? __ldcs(p.corrT + (int64_t)(c0 + i) * ext + a) : 0.f;
        }
        int32_t colr[16];
        if (fast) {
          // 16 independent cells in registers
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float4 wb = cw[i];
            float sim;
            if (kMode == 0) {   // both groups from TMEM: one packed evaluation
              const float2 cf = sg_coeff2(make_float2(__uint_as_float(uT[i]), __uint_as_float(uM[i])), waS, waQ,
                                          make_float2(wb.x, wb.y), make_float2(wb.z, wb.w), negInv);
              const float2 bl = __fmul2_rn(cf, make_float2(wT, wS));
              sim = __fadd_rn(bl.x, bl.y);
            } else {
              float temporal = 0.f, spectral = 0.f;
              if (useT) temporal = sg_coeff_h(__uint_as_float(uT[i]), wa.x, wa.z, wb.x, wb.z, invNT);
              if (kMode == 2 && p.loadT) temporal = tv[i];
              if (useS) spectral = sg_coeff_h(__uint_as_float(uM[i]), wa.y, wa.w, wb.y, wb.w, invNS);
              sim = __fadd_rn(__fmul_rn(temporal, wT), __fmul_rn(spectral, wS));
            }
            colr[i] = sg_grey(sim, b.colorScale, b.colorInv);
          }
          // the pixel (image row ext-1-c, x = a: coalesced over the lanes) and the transpose buffer
          const uint32_t offD = (uint32_t)(ext - 1 - c0) * (uint32_t)ext + (uint32_t)a;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            Pw[lane * kSgPPitch + i] = colr[i];
            b.rgb[offD - (uint32_t)i * (uint32_t)ext] = colr[i];
          }
          __syncwarp();
          // the mirrored pixels: image row ext-1-a', x = c: 16 consecutive pixels per row of this batch
          int32_t m[16];
#pragma unroll
          for (int it = 0; it < 16; it++) m[it] = Pw[(2 * it + (lane >> 4)) * kSgPPitch + (lane & 15)];
          const uint32_t offM = (uint32_t)(ext - 1 - (ta + quarter * 32 + (lane >> 4))) * (uint32_t)ext + (uint32_t)(c0 + (lane & 15));
#pragma unroll
          for (int it = 0; it < 16; it++) b.rgb[offM - (uint32_t)(2 * it) * (uint32_t)ext] = m[it];
          __syncwarp();
          continue;
        }
        // generic path: image edges, the diagonal, colour tables, colorWarp != 1, sim matrix output
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 wb = cw[i];
          float temporal = 0.f, spectral = 0.f;
          if (useT) temporal = sg_coeff_h(__uint_as_float(uT[i]), wa.x, wa.z, wb.x, wb.z, invNT);
          if (kMode == 2 && p.loadT) temporal = tv[i];
          if (useS) spectral = sg_coeff_h(__uint_as_float(uM[i]), wa.y, wa.w, wb.y, wb.w, invNS);
          const float sim = __fadd_rn(__fmul_rn(temporal, wT), __fmul_rn(spectral, wS));
          colr[i] = sg_color(b, sim, warpOne);
          if (p.simMat && rowOk && c0 + i < ext && c0 + i >= a) p.simMat[(int64_t)a * ext + c0 + i] = sim;
        }
        int32_t *row0 = b.rgb + (int64_t)(ext - 1 - c0) * ext + a;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          Pw[lane * kSgPPitch + i] = colr[i];
          if (rowOk && c0 + i < ext && c0 + i >= a) row0[-(int64_t)i * ext] = colr[i];   // upper triangle only
        }
        __syncwarp();
        {
          const int a2 = ta + quarter * 32 + (lane >> 4), c2 = c0 + (lane & 15);
          int32_t *mrow = b.rgb + (int64_t)(ext - 1 - a2) * ext + c2;
#pragma unroll 4
          for (int it = 0; it < 16; it++) {
            const int a3 = a2 + 2 * it;
            const int32_t mv = Pw[(2 * it + (lane >> 4)) * kSgPPitch + (lane & 15)];
            if (a3 < ext && a3 >= b.colBegin && a3 < b.colEnd && c2 < ext && c2 >= a3) mrow[-(int64_t)(2 * it) * ext] = mv;
          }
        }
        __syncwarp();
      }
      eMain += clock64() - tE;
    }
    if (p.prof && et == 0) {
      long long *o = p.prof + 16 * blockIdx.x + 12;
      o[0] = eAcc; o[1] = eMain;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace sgz
