// corr_tc.cuh -- K1 on the tensor cores (tcgen05 / TMEM): split-FP16 ("FP16x3") variant of corr_kernel.cuh.
//
//   cross(t) = sum_c sum_i q~[c][i] * b[c][t + i]            (FeatureCorrelationImpl.scala:198-210 via MathUtil.correlate)
//
// as a GEMM whose operands are never materialised in global memory:
//   A (M = 128, K-major, FP16, SWIZZLE_64B)  A[r][k] = b[c][t0 + 32 r + k]  -- a Hankel view of the channel row: the row
//        pitch of the swizzle atom (64 B = 32 halves) IS the shift between consecutive rows, a K step of 16 is +32 B, and
//        K blocks beyond one atom row simply run on into the next row (tools/umma_probe.cu: works with base_offset 0,
//        the hardware swizzles on absolute address bits);
//   B (N = 32, K-major, FP16, SWIZZLE_32B)   banded Toeplitz taps T[k][c'] = q~[k - (31 - c')], column order reversed so
//        that the 8x16 blocks of consecutive K steps alias each other: block (step s, row group g') = atom 2 s + g';
//   D[r][c'] = cross(t0 + 32 r + 31 - c'),  K = 32 + W  (+19 % MACs at W = 172), one tile = 4096 offsets.
// Precision (tools/tf32x3_probe.py, profiles/r01_tf32x3_precision.json): a single reduced-precision product misses the
// 1e-5 bar, so both operands are split into two FP16 numbers (22 significant bits, the data are normalised features of
// order 1) and   a*t = a1*t1 + (a2*t1 + a1*t2)   with the big term accumulated PER CHANNEL (chains of K/16 = 13 MMAs;
// the tensor core truncates when it aligns addends, long chains drift) and the small terms in a separate accumulator
// per group; the FP32 sums of the channel accumulators happen in the epilogue.  FP16 instead of TF32 halves the number of
// MMAs (K = 16 per instruction at 44 cycles vs K = 8 at 50 -- the A-operand fetch bounds small-N MMAs either way).
// 16 accumulators x 32 columns = all 512 TMEM columns: [T main, T corr, S corr, 13 spectral mains].
//
// The split warps, which touch every element anyway, also build the per-frame sums; the epilogue warps turn them into the
// window statistics (FP64 row sums + sliding, the FFMA kernel's arithmetic), the sim / boost curves and the file maxima.
#pragma once
#include <cuda_fp16.h>

#include "corr_kernel.cuh"

namespace sgz {

constexpr int kTcP = 32, kTcM = 128, kTcTile = kTcP * kTcM;
// the second FP16 parts (residuals, 2^-11 of the value) are stored times 2^11 so that they stay FP16 normals; their
// products go to separate accumulators which the epilogue scales back
constexpr float kTcLoScale = 2048.f;

struct TcGeom {
  int W, KS, natom, rows;
  uint32_t chanBytes;      // one FP16 operand buffer (first or second part) of one channel: rows x 64 B, multiple of 1024
  uint32_t tapsChanBytes;  // taps atoms of one channel: first then second FP16 part, natom x 256 B each
  uint32_t tapsPairBytes;  // two channels, rounded up to 1024
  uint32_t t0Bytes, fBytes, rsBytes;   // per-frame loudness / spectral sums, FP64 row sums
  size_t smemBytes;
};

__host__ __device__ inline TcGeom tc_geom(int W) {
  TcGeom g;
  g.W = W;
  g.KS = (kTcP + W + 15) / 16;
  g.natom = 2 * g.KS + 2;
  g.rows = kTcM + (g.KS * 16 + 31) / 32 + 1;
  g.chanBytes = (uint32_t)((g.rows * 64 + 1023) / 1024 * 1024);
  g.tapsChanBytes = (uint32_t)(g.natom * 256 * 2);
  g.tapsPairBytes = (uint32_t)((2 * g.tapsChanBytes + 1023) / 1024 * 1024);
  g.t0Bytes = (uint32_t)((g.rows * 33 * 4 + 127) / 128 * 128);   // skewed: frame e at index e + e / 32
  g.fBytes = (uint32_t)((g.rows * 33 * 8 + 127) / 128 * 128);
  g.rsBytes = (uint32_t)((g.rows * 5 + g.rows * 17) * 8 + 64);   // FP64 sums of whole rows (stride 5) + of 8-frame quarters (stride 17): padded against bank conflicts
  g.smemBytes = (size_t)8 * g.chanBytes + (size_t)2 * g.tapsPairBytes + g.t0Bytes + g.fBytes + g.rsBytes +
                1024 /*alignment slack*/ + 256 /*barriers*/;
  return g;
}

// taps of all channel pairs in the exact shared-memory image the kernel bulk-copies: per pair [chan x][chan y],
// per channel [first-part atoms][second-part atoms], atom a = 8 rows (cc) x 16 k (kk) halves with the SWIZZLE_32B
// chunk flip on rows 4..7
inline void tc_build_taps(const std::vector<float> &pairTaps /*[numPairs][Wq] float2*/, int numPairs, int Wq, int W,
                          std::vector<uint16_t> &out) {
  const TcGeom g = tc_geom(W);
  out.assign((size_t)numPairs * g.tapsPairBytes / 2, 0);
  for (int p = 0; p < numPairs; p++)
    for (int h = 0; h < 2; h++)
      for (int part = 0; part < 2; part++)
        for (int a = 0; a < g.natom; a++)
          for (int cc = 0; cc < 8; cc++)
            for (int kk = 0; kk < 16; kk++) {
              const int q = 8 * a + kk + cc - 31;
              __half v = __float2half_rn(0.f);
              if (q >= 0 && q < W) {
                const float tp = pairTaps[((size_t)p * Wq + q) * 2 + h];
                const __half t1 = __float2half_rn(tp);
                v = part == 0 ? t1 : __float2half_rn((tp - __half2float(t1)) * kTcLoScale);
              }
              const size_t byteOff = (size_t)p * g.tapsPairBytes + (size_t)h * g.tapsChanBytes +
                                     (size_t)part * g.natom * 256 + (size_t)a * 256 + (size_t)cc * 32 +
                                     (size_t)(((kk >> 3) ^ ((cc >> 2) & 1)) << 4) + (size_t)(kk & 7) * 2;
              uint16_t bits;
              memcpy(&bits, &v, 2);
              out[byteOff / 2] = bits;
            }
}

// ---------------------------------------------------------------------------------------------
// the kernel: cross terms on the tensor cores, window statistics and the final sim fused
// ---------------------------------------------------------------------------------------------
struct CorrTcParams {
  const float2 *data;
  int64_t rowStride, usedFrames;
  int numCh, numPairs, W;
  const uint16_t *taps;     // tc_build_taps image, numPairs x tapsPairBytes
  double stdT, stdS, rhoT, rhoS, lnAvgIn;
  float weight, maxBoost;
  const int64_t *fileStart;
  const int32_t *tileFile;      // [numTiles + 1] file that holds frame 4096 * tile (clamped to the last file)
  int numFiles, tailExtra;
  int64_t tileBegin, tileEnd;   // tiles of 4096 offsets
  float *sim, *boost;
  unsigned long long *fileMax;  // [numFiles] packed (order_key(sim) << 32 | ~offset), or nullptr
  long long *prof;              // developer probe (SGZ_CORR_TC_PROF): per CTA 16 cycle counters, or nullptr
  int chainOrder;               // developer knob (SGZ_CORR_TC_CHAIN=1): one chain per product instead of the per-K-step order
};

__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ void tc_mma(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmemD),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// A-operand collector: `fill` keeps the fetched A in the tensor core's operand buffer, `lastuse` takes it from there
// instead of shared memory (SASS: UTCHMMA ... .A_KEEP / .A_REUSE)
__device__ __forceinline__ void tc_mma_fill(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmemD),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_lastuse(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmemD),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// bounded wait: a protocol error must end in a trap (visible CUDA error), not in a hung GPU box
template <bool kSleep = true>
__device__ __forceinline__ void tc_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 24) && !done; it++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (kSleep && !done) __nanosleep(100);
  }
  if (!done) __trap();
}
__device__ __forceinline__ bool tc_elect() {   // one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// warps 0..6: split workers (global pair row -> swizzled FP16 operand buffers + per-frame sums), warp 7: MMA issuer
// (one elected lane), warps 8..15: epilogue (TMEM lane quarter = warp % 4; window statistics, sim, file maxima)
// 7 + 1 + 8 = 16 warps: the register file is handed out to groups of 4 warps, 17 warps would cost 128 -> 96 registers
constexpr int kTcSplit = 224, kTcEpi = 256, kTcThreads = kTcSplit + 32 + kTcEpi;

__global__ void __launch_bounds__(kTcThreads, 1) k_corr_tc(const CorrTcParams p) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  const TcGeom G = tc_geom(p.W);
  unsigned char *base = smemRaw + ((1024 - (smem_u32(smemRaw) & 1023)) & 1023);
  // operand buffers: [buf 0/1][channel x/y][first / second FP16 part]
  auto ops = [&](int buf, int h, int part) { return base + (size_t)((buf * 2 + h) * 2 + part) * G.chanBytes; };
  unsigned char *tapsBase = base + (size_t)8 * G.chanBytes;
  auto tapsBuf = [&](int buf) { return tapsBase + (size_t)buf * G.tapsPairBytes; };
  float *T0 = reinterpret_cast<float *>(tapsBase + (size_t)2 * G.tapsPairBytes);           // loudness per frame
  float2 *F = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(T0) + G.t0Bytes);   // (sum_c b, sum_c b^2)
  double *RS = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(F) + G.fBytes);    // [rows][4] row sums
  double *RQ = RS + 5 * G.rows;                                                                // [rows][4 quarters][4] + 1 pad
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(RS) + G.rsBytes);
  uint64_t *opFree = bars, *opFull = bars + 2, *tapsFull = bars + 4, *accFull = bars + 6, *accEmpty = bars + 7;
  uint64_t *statsFull = bars + 8, *statsFree = bars + 9;
  uint32_t *tmemSlot = reinterpret_cast<uint32_t *>(bars + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(opFree, 1); mbar_init(opFree + 1, 1);
    mbar_init(opFull, kTcSplit / 32); mbar_init(opFull + 1, kTcSplit / 32);
    mbar_init(tapsFull, 1); mbar_init(tapsFull + 1, 1);
    mbar_init(accFull, 1);
    mbar_init(accEmpty, kTcEpi / 32);
    mbar_init(statsFull, kTcSplit / 32);
    mbar_init(statsFree, kTcEpi / 32);
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
  const int nFrames = G.rows * 32;                      // frames of a tile that the A operand can touch

  if (warp < kTcSplit / 32) {
    // =========================== split workers ===========================
    constexpr int kPer = 10;                            // >= ceil(nFrames / 448) for W <= 256; two frames per step
    uint32_t pc = 0, tileIt = 0;
    long long sLoad = 0, sWait = 0, sStore = 0, sStats = 0, tS;
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, tileIt++) {
      const int64_t t0 = tile * kTcTile;
      float b0[2 * kPer], s1[2 * kPer], s2[2 * kPer];   // per-frame sums over the channels of this thread's frames
      for (int pr = 0; pr < p.numPairs; pr++, pc++) {
        const int buf = pc & 1;
        const uint32_t use = pc >> 1;
        const float4 *row = reinterpret_cast<const float4 *>(p.data + (int64_t)pr * p.rowStride + t0);   // t0 % 4096 == 0
        unsigned char *x1 = ops(buf, 0, 0), *x2 = ops(buf, 0, 1), *y1 = ops(buf, 1, 0), *y2 = ops(buf, 1, 1);
        tS = clock64();
        {   // pull the NEXT pair row into L2 now: its loads then see an L2 hit instead of the HBM latency
          const bool lastPair = pr == p.numPairs - 1;
          const int64_t nt0 = lastPair ? t0 + (int64_t)gridDim.x * kTcTile : t0;
          if (!lastPair || tile + gridDim.x < p.tileEnd) {
            const float4 *nrow = reinterpret_cast<const float4 *>(p.data + (int64_t)(lastPair ? 0 : pr + 1) * p.rowStride + nt0);
#pragma unroll
            for (int k = 0; k < kPer; k++) {
              const int L2 = tid + k * kTcSplit;
              if ((L2 & 7) == 0 && 2 * L2 < nFrames) asm volatile("prefetch.global.L2 [%0];" ::"l"(nrow + L2));
            }
          }
        }
#pragma unroll
        for (int k0 = 0; k0 < kPer; k0 += kPer) {       // all loads of the pair row in flight at once
          float4 v[kPer];
#pragma unroll
          for (int k = 0; k < kPer; k++) {
            const int L2 = tid + (k0 + k) * kTcSplit;   // frames 2 L2, 2 L2 + 1
            v[k] = 2 * L2 < nFrames ? __ldg(row + L2) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (pr == 0) {   // pair 0 = (loudness, first spectral channel)
#pragma unroll
            for (int k = 0; k < kPer; k++) {
              const int kk = k0 + k;
              b0[2 * kk] = v[k].x; s1[2 * kk] = v[k].y; s2[2 * kk] = v[k].y * v[k].y;
              b0[2 * kk + 1] = v[k].z; s1[2 * kk + 1] = v[k].w; s2[2 * kk + 1] = v[k].w * v[k].w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < kPer; k++) {
              const int kk = k0 + k;
              s1[2 * kk] += v[k].x + v[k].y; s2[2 * kk] = fmaf(v[k].x, v[k].x, fmaf(v[k].y, v[k].y, s2[2 * kk]));
              s1[2 * kk + 1] += v[k].z + v[k].w; s2[2 * kk + 1] = fmaf(v[k].z, v[k].z, fmaf(v[k].w, v[k].w, s2[2 * kk + 1]));
            }
          }
          sLoad += clock64() - tS; tS = clock64();
          if (k0 == 0 && use > 0) tc_wait<false>(opFree + buf, (use - 1) & 1);   // MMAs that last used this buffer are done
          sWait += clock64() - tS; tS = clock64();
#pragma unroll
          for (int k = 0; k < kPer; k++) {
            const int L = 2 * (tid + (k0 + k) * kTcSplit);
            if (L < nFrames) {
              // frame L -> row L / 32 (64 B), 16-byte chunk (L % 32) / 8 flipped by (row >> 1) & 3, half (L % 8)
              const int r = L >> 5, ch = (L & 31) >> 3, w = L & 7;
              const uint32_t off = (uint32_t)(r * 64 + ((ch ^ ((r >> 1) & 3)) << 4) + w * 2);
              const __half2 xa = __floats2half2_rn(v[k].x, v[k].z), ya = __floats2half2_rn(v[k].y, v[k].w);
              const float2 xf = __half22float2(xa), yf = __half22float2(ya);
              *reinterpret_cast<__half2 *>(x1 + off) = xa;
              *reinterpret_cast<__half2 *>(x2 + off) = __floats2half2_rn((v[k].x - xf.x) * kTcLoScale, (v[k].z - xf.y) * kTcLoScale);
              *reinterpret_cast<__half2 *>(y1 + off) = ya;
              *reinterpret_cast<__half2 *>(y2 + off) = __floats2half2_rn((v[k].y - yf.x) * kTcLoScale, (v[k].w - yf.y) * kTcLoScale);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(opFull + buf);
        sStore += clock64() - tS;
      }
      tS = clock64();
      // per-frame sums of the tile -> shared memory for the epilogue warps
      if (tileIt > 0) tc_wait(statsFree, (tileIt - 1) & 1);
#pragma unroll
      for (int k = 0; k < kPer; k++) {
        const int L = 2 * (tid + k * kTcSplit);
        if (L < nFrames) {   // skewed by one element per 32-frame row: the epilogue lanes walk rows, not frames
          const int i = L + (L >> 5);
          T0[i] = b0[2 * k]; T0[i + 1] = b0[2 * k + 1];
          F[i] = make_float2(s1[2 * k], s2[2 * k]); F[i + 1] = make_float2(s1[2 * k + 1], s2[2 * k + 1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(statsFull);
      sStats += clock64() - tS;
    }
    if (p.prof && tid == 0) {
      long long *o = p.prof + 24 * blockIdx.x + 16;
      o[0] = sLoad; o[1] = sWait; o[2] = sStore; o[3] = sStats;
    }
  } else if (warp == kTcSplit / 32) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the (uniform) control flow and one elected lane issues: descriptors and loop counters then
    // live in uniform registers and each tcgen05.mma costs a handful of instructions.
    // D = F32, A = B = F16 (format 0), both K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kTcP >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
    uint32_t pc = 0, tileIt = 0;
    long long cFree = 0, cAcc = 0, cFull = 0, cTaps = 0, cIssue = 0, cTotal = clock64(), tA;
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, tileIt++) {
      bool corrStartedT = false, corrStartedS = false;
      for (int pr = 0; pr < p.numPairs; pr++, pc++) {
        const int buf = pc & 1;
        const uint32_t use = pc >> 1;
        if (pc == 0 && tc_elect()) {   // very first pair: nothing to wait for (later requests are issued one pair ahead)
          mbar_expect_tx(tapsFull + buf, G.tapsPairBytes);
          bulk_g2s(tapsBuf(buf), p.taps, G.tapsPairBytes, tapsFull + buf);
        }
        tA = clock64();
        if (pr == 0 && tileIt > 0) tc_wait<false>(accEmpty, (tileIt - 1) & 1);   // epilogue has drained the accumulators
        cAcc += clock64() - tA;
        tA = clock64();
        tc_wait<false>(opFull + buf, use & 1);
        cFull += clock64() - tA;
        tA = clock64();
        tc_wait<false>(tapsFull + buf, use & 1);
        cTaps += clock64() - tA;
        tA = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int h = 0; h < 2; h++) {
          const int c = 2 * pr + h;
          if (c >= p.numCh) break;
          const uint32_t dMain = tmem + 32u * (c == 0 ? 0u : (uint32_t)(2 + c));
          const uint32_t dCorr = tmem + 32u * (c == 0 ? 1u : 2u);
          const uint32_t tHiA = smem_u32(tapsBuf(buf)) + (uint32_t)h * G.tapsChanBytes;
          const uint64_t aHi = tc_desc(smem_u32(ops(buf, h, 0)), 512, 4), aLo = tc_desc(smem_u32(ops(buf, h, 1)), 512, 4);
          const uint64_t tHi = tc_desc(tHiA, 256, 6), tLo = tc_desc(tHiA + (uint32_t)G.natom * 256u, 256, 6);
          const bool started = c == 0 ? corrStartedT : corrStartedS;
          if (c == 0) corrStartedT = true; else corrStartedS = true;
          if (tc_elect()) {
            // one K step of 16: A start +32 B (2 descriptor units), taps start +2 atoms = 512 B (32 units).
            // N = 32 MMAs are bound by the fetch of the 4 KB A operand: per K step a1 t1 and a1 t2 share A through the
            // operand collector (.collector::a::fill / lastuse), so two of three MMAs fetch it (changing the
            // accumulator between MMAs costs nothing, tools/umma_rate_probe.cu).  SGZ_CORR_TC_CHAIN=1: the round-1 order.
            if (p.chainOrder) {
              uint64_t da = aHi, db = tHi;
              for (int s = 0; s < G.KS; s++, da += 2, db += 32) tc_mma(dMain, da, db, idesc, s > 0);
              da = aLo; db = tHi;
              for (int s = 0; s < G.KS; s++, da += 2, db += 32) tc_mma(dCorr, da, db, idesc, started || s > 0);
              da = aHi; db = tLo;
              for (int s = 0; s < G.KS; s++, da += 2, db += 32) tc_mma(dCorr, da, db, idesc, 1);
            } else {
              uint64_t d1 = aHi, d2 = aLo, b1 = tHi, b2 = tLo;
              for (int s = 0; s < G.KS; s++, d1 += 2, d2 += 2, b1 += 32, b2 += 32) {
                tc_mma_fill(dMain, d1, b1, idesc, s > 0);
                tc_mma_lastuse(dCorr, d1, b2, idesc, started || s > 0);
                tc_mma(dCorr, d2, b1, idesc, 1);
              }
            }
          }
        }
        if (tc_elect()) {
          tc_commit(opFree + buf);
          if (pr == p.numPairs - 1) tc_commit(accFull);
        }
        __syncwarp();
        cIssue += clock64() - tA;
        // taps of the NEXT pair, one pair ahead: its buffer is free once the pair before this one has executed, which
        // (tcgen05 ops run in issue order) happens while the MMAs just issued are still queued
        {
          const bool lastPair = pr == p.numPairs - 1;
          if (!(lastPair && tile + gridDim.x >= p.tileEnd)) {
            const uint32_t npc = pc + 1, nbuf = npc & 1, nuse = npc >> 1;
            const int npr = lastPair ? 0 : pr + 1;
            tA = clock64();
            if (nuse > 0) tc_wait<false>(opFree + nbuf, (nuse - 1) & 1);
            cFree += clock64() - tA;
            if (tc_elect()) {
              mbar_expect_tx(tapsFull + nbuf, G.tapsPairBytes);
              bulk_g2s(tapsBuf(nbuf), reinterpret_cast<const unsigned char *>(p.taps) + (size_t)npr * G.tapsPairBytes,
                       G.tapsPairBytes, tapsFull + nbuf);
            }
          }
        }
      }
    }
    if (p.prof && lane == 0) {
      long long *o = p.prof + 24 * blockIdx.x;
      o[0] = clock64() - cTotal; o[1] = cFree; o[2] = cAcc; o[3] = cFull; o[4] = cTaps; o[5] = cIssue; o[6] = tileIt;
    }
  } else {
    // =========================== epilogue ===========================
    // 8 warps: TMEM lane quarter q = warp % 4 (rows r = 32 q + lane), column half hs = (warp - 8) / 4.  A thread owns the
    // 16 offsets t0 + 32 r + jb .. + 15 (columns 31 - offset), jb = 16 (1 - hs).  Two warps per scheduler: the FP64 /
    // MUFU latencies of the window statistics need the second warp to hide behind.
    const int ew = warp - (kTcSplit / 32 + 1), quarter = warp & 3, hs = ew >> 2, et = ew * 32 + lane;   // et = 0..255
    const int jb = 16 * (1 - hs);
    const int W = p.W;
    const bool useT = p.weight > 0.f, useS = p.weight < 1.f;
    const float qnan = __int_as_float(0x7fc00000);
    const double invW = 1.0 / (double)W, invNS = 1.0 / ((double)(p.numCh - 1) * (double)W);
    const float cT = (float)(invW / p.stdT), cS = (float)(invNS / p.stdS);
    const float rhoT = (float)p.rhoT, rhoS = (float)p.rhoS, l2In = (float)(p.lnAvgIn * 1.4426950408889634);
    uint32_t tileIt = 0;
    long long eAcc = 0, eLd = 0, eSt = 0, eMain = 0, eRow = 0, eInit = 0;
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, tileIt++) {
      const int64_t t0 = tile * kTcTile;
      // file of this thread's first offset: the tile spans files [tileFile[tile], tileFile[tile + 1]] (host table)
      const int r = quarter * 32 + lane;
      const int64_t g0 = t0 + 32 * (int64_t)r + jb;
      int f;
      {
        int lo = p.tileFile[tile], hi = p.tileFile[tile + 1] + 1;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (p.fileStart[mid] <= g0) lo = mid; else hi = mid;
        }
        f = lo;
      }
      int64_t fStart = p.fileStart[f], fEnd = p.fileStart[f + 1];
      long long tE = clock64();
      tc_wait(accFull, tileIt & 1);
      eAcc += clock64() - tE; tE = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t laneAddr = tmem + ((uint32_t)(quarter * 32) << 16) + 16u * (uint32_t)hs;
      float accT[16], accS[16];
      {
        uint32_t u[16], w[16];
        tc_ld16_nowait(laneAddr + 0 * 32, u);
        tc_ld16_nowait(laneAddr + 1 * 32, w);
        tc_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) accT[i] = fmaf(__uint_as_float(w[i]), 1.0f / kTcLoScale, __uint_as_float(u[i]));
        tc_ld16_nowait(laneAddr + 2 * 32, u);
        tc_ld16_nowait(laneAddr + 3 * 32, w);
        tc_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) accS[i] = fmaf(__uint_as_float(u[i]), 1.0f / kTcLoScale, __uint_as_float(w[i]));
        for (int c = 2; c < p.numCh; c += 2) {     // two accumulators per wait; an odd tail reads one
          tc_ld16_nowait(laneAddr + (uint32_t)(2 + c) * 32, u);
          if (c + 1 < p.numCh) tc_ld16_nowait(laneAddr + (uint32_t)(3 + c) * 32, w);
          tc_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i++) accS[i] += __uint_as_float(u[i]) + (c + 1 < p.numCh ? __uint_as_float(w[i]) : 0.f);
        }
      }
      // the accumulators are in registers: hand TMEM back to the issuer
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(accEmpty);
      eLd += clock64() - tE; tE = clock64();

      // ---- window statistics: FP64 sums of whole 32-frame rows, then slide frame by frame ----
      tc_wait(statsFull, tileIt & 1);
      eSt += clock64() - tE; tE = clock64();
      for (int qd = et; qd < 4 * G.rows; qd += kTcEpi) {      // FP64 sums of every 8-frame quarter row
        const int e0 = 33 * (qd >> 2) + 8 * (qd & 3);
        double x[8], a1, a2, a3, a4;
        float2 fq[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { x[k] = (double)T0[e0 + k]; fq[k] = F[e0 + k]; }
        a1 = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
        a2 = ((x[0] * x[0] + x[1] * x[1]) + (x[2] * x[2] + x[3] * x[3])) + ((x[4] * x[4] + x[5] * x[5]) + (x[6] * x[6] + x[7] * x[7]));
        a3 = (((double)fq[0].x + (double)fq[1].x) + ((double)fq[2].x + (double)fq[3].x)) +
             (((double)fq[4].x + (double)fq[5].x) + ((double)fq[6].x + (double)fq[7].x));
        a4 = (((double)fq[0].y + (double)fq[1].y) + ((double)fq[2].y + (double)fq[3].y)) +
             (((double)fq[4].y + (double)fq[5].y) + ((double)fq[6].y + (double)fq[7].y));
        double *o = RQ + 17 * (qd >> 2) + 4 * (qd & 3);
        o[0] = a1; o[1] = a2; o[2] = a3; o[3] = a4;
      }
      tc_epi_sync();
      if (et < G.rows) {                                      // whole rows from their quarters
        const double *q = RQ + 17 * et;
        double *o = RS + 5 * et;
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = (q[k] + q[4 + k]) + (q[8 + k] + q[12 + k]);
      }
      tc_epi_sync();
      eRow += clock64() - tE; tE = clock64();
      // window of the first offset: frames [32 r + jb, 32 r + jb + W) = quarters up to a row boundary, whole rows,
      // quarters, and at most 7 single frames (jb is 0 or 16, so the start is quarter aligned)
      D4 win = {0, 0, 0, 0};
      {
        int left = W, fr = 32 * r + jb;                       // fr = next frame (relative to the tile) to add
        while ((fr & 31) != 0 && left >= 8) {
          const double *o = RQ + 17 * (fr >> 5) + 4 * ((fr >> 3) & 3);
          win.t1 += o[0]; win.t2 += o[1]; win.s1 += o[2]; win.s2 += o[3];
          fr += 8; left -= 8;
        }
        D4 w2 = {0, 0, 0, 0};                                 // second chain: halves the dependent FP64 adds
        for (; left >= 64; left -= 64, fr += 64) {
          const double *o = RS + 5 * (fr >> 5);
          win.t1 += o[0]; win.t2 += o[1]; win.s1 += o[2]; win.s2 += o[3];
          w2.t1 += o[5]; w2.t2 += o[6]; w2.s1 += o[7]; w2.s2 += o[8];
        }
        if (left >= 32) {
          const double *o = RS + 5 * (fr >> 5);
          win.t1 += o[0]; win.t2 += o[1]; win.s1 += o[2]; win.s2 += o[3];
          fr += 32; left -= 32;
        }
        while (left >= 8) {
          const double *o = RQ + 17 * (fr >> 5) + 4 * ((fr >> 3) & 3);
          w2.t1 += o[0]; w2.t2 += o[1]; w2.s1 += o[2]; w2.s2 += o[3];
          fr += 8; left -= 8;
        }
        const int e0 = 33 * (fr >> 5) + (fr & 31);
#pragma unroll
        for (int k = 0; k < 7; k++) {
          if (k < left) {
            const double x = (double)T0[e0 + k];
            const float2 fq = F[e0 + k];
            w2.t1 += x; w2.t2 += x * x; w2.s1 += (double)fq.x; w2.s2 += (double)fq.y;
          }
        }
        win.t1 += w2.t1; win.t2 += w2.t2; win.s1 += w2.s1; win.s2 += w2.s2;
      }
      eInit += clock64() - tE; tE = clock64();
      unsigned long long best = 0ull;
      const float *T0r = T0 + 33 * r;      // row base in the skewed arrays: frame 32 r + x sits at x + x / 32
      const float2 *Fr = F + 33 * r;
      // common case: all 16 offsets are evaluated offsets of ONE file -> no control flow per offset
      const bool plain = g0 + 16 <= fStart + ((fEnd - fStart) - p.tailExtra - W + 1) && g0 + 16 <= p.usedFrames;
      const uint32_t tl0 = (uint32_t)(g0 - fStart);
#pragma unroll
      for (int blk = 0; blk < 2; blk++) {
        // pass A: straight-line arithmetic for 8 offsets, so that the 8 dependency chains overlap
        float simv[8], boostv[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const int jj = 8 * blk + e;           // offset inside the thread's run; accumulator column 15 - jj of its half
          const double mT = win.t1 * invW;
          const float avgB = (float)mT;                                          // MathUtil.avg -> Float
          const float boost = exp2f((l2In - __log2f(avgB)) * (1.0f / 0.6f));     // calcBoost
          float temporal = 0.f, spectral = 0.f;
          if (useT) {
            const double q = win.t2 * invW;
            const double var = q - mT * mT;
            const float cr = accT[15 - jj] - (float)mT * rhoT;
            temporal = (var > 1e-13 * q) ? (cr * cT) * rsqrtf((float)var) : qnan;
          }
          if (useS) {
            const double mS = win.s1 * invNS;
            const double q = win.s2 * invNS;
            const double var = q - mS * mS;
            const float cr = accS[15 - jj] - (float)mS * rhoS;
            spectral = (var > 1e-13 * q) ? (cr * cS) * rsqrtf((float)var) : qnan;
          }
          const float blend = __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
          simv[e] = boost <= p.maxBoost ? blend : 0.f;
          boostv[e] = boost;
          if (jj < 15) {
            // slide the window by one frame; exact FP64 deltas: a window that has moved onto constant data (digital
            // silence) must end up with EXACTLY zero variance, the reference yields NaN there
            const int jo = jb + jj, jn = jo + W;
            const int io = jo + (jo >> 5), in = jn + (jn >> 5);
            const double bo = (double)T0r[io], bn = (double)T0r[in];
            const float2 fo = Fr[io], fn = Fr[in];
            win.t1 += bn - bo;
            win.t2 += bn * bn - bo * bo;
            win.s1 += (double)fn.x - (double)fo.x;
            win.s2 += (double)fn.y - (double)fo.y;
          }
        }
        // pass B: which offsets exist (window inside its file), per-file maximum (first occurrence)
        if (plain) {
#pragma unroll
          for (int e = 0; e < 8; e++) {
            if (simv[e] == simv[e]) {
              const unsigned long long key = ((unsigned long long)float_order_key(simv[e]) << 32) |
                                             (unsigned long long)(0xffffffffu - (tl0 + (uint32_t)(8 * blk + e)));
              if (key > best) best = key;
            }
          }
        } else {
          for (int e = 0; e < 8; e++) {
            const int64_t g = g0 + 8 * blk + e;
            while (g >= fEnd && f + 1 < p.numFiles) {
              if (best != 0ull && p.fileMax) atomicMax(p.fileMax + f, best);
              best = 0ull;
              f++;
              fStart = fEnd;
              fEnd = p.fileStart[f + 1];
            }
            const int64_t tl = g - fStart;
            float sv = qnan, bv = qnan;
#pragma unroll
            for (int kk = 0; kk < 8; kk++) if (kk == e) { sv = simv[kk]; bv = boostv[kk]; }
            if (!(g < p.usedFrames && tl < (fEnd - fStart) - p.tailExtra - W + 1)) { sv = qnan; bv = qnan; }
            else if (sv == sv) {
              const unsigned long long key = ((unsigned long long)float_order_key(sv) << 32) |
                                             (unsigned long long)(0xffffffffu - (uint32_t)tl);
              if (key > best) best = key;
            }
#pragma unroll
            for (int kk = 0; kk < 8; kk++) if (kk == e) { simv[kk] = sv; boostv[kk] = bv; }
          }
        }
        float4 *so = reinterpret_cast<float4 *>(p.sim + g0 + 8 * blk), *bo = reinterpret_cast<float4 *>(p.boost + g0 + 8 * blk);
        so[0] = make_float4(simv[0], simv[1], simv[2], simv[3]);
        so[1] = make_float4(simv[4], simv[5], simv[6], simv[7]);
        bo[0] = make_float4(boostv[0], boostv[1], boostv[2], boostv[3]);
        bo[1] = make_float4(boostv[4], boostv[5], boostv[6], boostv[7]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(statsFree);   // T0 / F / RS may be rebuilt
      if (best != 0ull && p.fileMax) atomicMax(p.fileMax + f, best);
      eMain += clock64() - tE;
    }
    if (p.prof && et == 0) {
      long long *o = p.prof + 24 * blockIdx.x + 8;
      o[0] = eAcc; o[1] = eLd; o[2] = eSt; o[3] = eMain; o[4] = eRow; o[5] = eInit;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace sgz
