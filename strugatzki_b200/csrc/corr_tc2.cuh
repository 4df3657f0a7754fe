// corr_tc2.cuh -- K1 on the tensor cores, second generation: N = 64 tiles fed by bulk copies (round 2).
//
//   cross(t) = sum_c sum_i q~[c][i] * b[c][t + i]            (FeatureCorrelationImpl.scala:198-210 via MathUtil.correlate)
//
// Same split-FP16 formulation as corr_tc.cuh (a*t = a1*t1 + (a2*t1 + a1*t2), products accumulated in FP32 in TMEM), but
//   * the Hankel period is 64: A (M = 128, K-major, SWIZZLE_128B) A[r][k] = b[c][t0 + 64 r + k] -- the 128-byte row of the
//     swizzle atom IS the shift between rows, so the operand is the FP16 signal itself and one tile is 8192 offsets;
//     B (N = 64, K-major, SWIZZLE_32B) = banded Toeplitz taps with reversed columns, block (K step s, row group g') =
//     atom 2 s + g'.  An M128 x N64 x K16 MMA fetches 6 KB from shared memory for 64 columns where N = 32 fetched 5 KB
//     for 32 (the operand fetch, not the math, bounds small-N MMAs: 48 against 44.5 cycles, tools/umma_rate_probe.cu);
//   * the FP16 parts of the database are computed ONCE per database (k_db_planes: pre-swizzled planes, 2 x 2 B per value
//     = the 56 B per frame of the float rows) together with the per-frame sums the window statistics need, so nothing in
//     this kernel touches the operands with ordinary loads/stores: a producer lane streams them with cp.async.bulk
//     through a 4-deep ring (UBLKCP), the taps of the channel through a second ring;
//   * 8 accumulators of 64 columns fill TMEM: [T main, T corr, S corr, 5 spectral mains] -- the spectral channels take
//     turns on the five main accumulators (chains of <= 3 x KS MMAs; the tensor core truncates when it aligns addends,
//     so the chain length bounds the bias);
//   * 16 epilogue warps (window statistics in FP64 from 16-frame block sums + exact slides, boost, blend, file maxima)
//     work on tile n while the MMAs of tile n + 1 run; the FP64 block sums are part of the prepared database too and
//     arrive by bulk copy.
#pragma once
#include <cuda_fp16.h>

#include "corr_fix.cuh"
#include "corr_tc.cuh"

namespace sgz {

constexpr int kT2P = 64, kT2M = 128, kT2Tile = kT2P * kT2M;
constexpr int kT2Mains = 5;               // spectral main accumulators
constexpr int kT2Ahead = 0;               // channels between an L2 prefetch of a signal stage and its bulk copy: 0 = none (measured: no gain, the ring itself keeps the bytes in flight)
constexpr int kT2EpiWarps = 16, kT2Threads = (4 + kT2EpiWarps) * 32;

constexpr int kT2PassKS = 16;             // K steps per pass of a long window (chains of <= 3 x 16 MMAs per accumulator)
constexpr int kT2SinglePassKS = 20;       // windows up to 257 frames take one pass

struct T2Geom {
  int W, KS, natom, rows, sumRows, sumPitch;
  // long windows (more than kT2SinglePassKS K steps) are cut into NP passes of KS (= kT2PassKS) K steps, the last one of
  // KSlast: pass j multiplies the signal shifted by 256 j frames with the taps atoms [32 j, 32 j + natom); the epilogue
  // adds the accumulators of the passes in registers.  One pass: NP = 1, KS = KSlast = all K steps.
  int NP, KSlast, natomFull;
  uint32_t planeBytes;      // one FP16 part of one channel of one tile (rows x 128 B)
  uint32_t planeStride;     // placement stride in shared memory (multiple of 1024)
  uint32_t tapsBytes;       // taps of one channel and pass: first-part atoms then second-part atoms
  uint32_t tapsFullBytes;   // taps image of one channel in global memory: [natomFull first-part atoms][... second-part atoms]
  uint32_t tapsStride;
  uint32_t sumsBytes;
  int sigStages, tapStages;
  size_t smemBytes;
};

__host__ __device__ inline T2Geom t2_geom(int W, size_t smemMax = 232448, int ring = 0x43) {
  T2Geom g;
  g.W = W;
  const int ksAll = (kT2P - 1 + W + 15) / 16;
  if (ksAll <= kT2SinglePassKS) { g.KS = ksAll; g.NP = 1; g.KSlast = ksAll; }
  else { g.KS = kT2PassKS; g.NP = (ksAll + kT2PassKS - 1) / kT2PassKS; g.KSlast = ksAll - (g.NP - 1) * kT2PassKS; }
  g.natom = 2 * g.KS + 6;
  g.natomFull = 2 * g.KS * g.NP + 6;
  g.rows = kT2M - 1 + (g.KS + 3) / 4;               // row 127 runs on for 16 KS halves
  g.planeBytes = (uint32_t)g.rows * 128u;
  // odd passes start 4 rows into the 8-row swizzle period of their source: they are placed 512 bytes into the stage
  g.planeStride = (g.planeBytes + (g.NP > 1 ? 512u : 0u) + 1023u) / 1024u * 1024u;
  g.tapsBytes = (uint32_t)g.natom * 256u * 2u;
  g.tapsFullBytes = (uint32_t)g.natomFull * 256u * 2u;
  g.tapsStride = (g.tapsBytes + 1023u) / 1024u * 1024u;
  g.sumRows = kT2M + (W + 63) / 64;                  // 64-frame rows that windows starting in the tile can touch
  g.sumPitch = (g.sumRows + 1) & ~1;                 // bulk copies move multiples of 16 bytes
  g.sumsBytes = (uint32_t)(16 * g.sumPitch * 8);     // [4 quantities][4 column blocks][sumPitch] doubles
  g.sigStages = ring >> 4;          // initial ring depths (signal, taps); reduced below until they fit
  g.tapStages = ring & 15;
  for (;;) {
    g.smemBytes = (size_t)g.sigStages * 2 * g.planeStride + (size_t)g.tapStages * g.tapsStride + g.sumsBytes +
                  1024 /*alignment slack*/ + 256 /*barriers*/;
    if (g.smemBytes <= smemMax || (g.sigStages == 2 && g.tapStages == 2)) break;
    if (g.sigStages > g.tapStages || g.tapStages == 2) g.sigStages--; else g.tapStages--;
  }
  return g;
}

// position of frame g in a pre-swizzled FP16 plane (bytes): SWIZZLE_128B = 16-byte chunk index ^= 128-byte row index mod 8
__host__ __device__ inline int64_t t2_plane_byte(int64_t g) {
  const int64_t a = 2 * g;
  return a ^ (((a >> 7) & 7) << 4);
}
// position of frame g in a tile-transposed per-frame array: [tile][frame % 64][frame / 64 % 128]
__host__ __device__ inline int64_t t2_side_index(int64_t g) {
  return (g & ~(int64_t)(kT2Tile - 1)) + ((g & 63) << 7) + ((g >> 6) & 127);
}

// taps image of a taps stage: per channel [first-part atoms][second-part atoms], atom a = 8 rows (cc) x 16 k (kk) halves
// holding q~[8 a + kk + cc - 63], SWIZZLE_32B chunk flip on rows 4..7.  Built on the device from the float taps (one block
// per channel): the float taps are a few KB and reach the device as an inline copy even while database uploads occupy
// the copy engine (streaming scans)
__global__ void k_t2_taps(const float2 *__restrict__ pairTaps, int numCh, int Wq, int W, unsigned char *__restrict__ out) {
  const T2Geom g = t2_geom(W);   // natomFull / tapsFullBytes do not depend on the shared-memory limit
  const int c = blockIdx.x;
  const float *tp = reinterpret_cast<const float *>(pairTaps + (size_t)(c >> 1) * Wq) + (c & 1);
  for (int i = threadIdx.x; i < 2 * g.natomFull * 128; i += blockDim.x) {
    const int part = i / (g.natomFull * 128), rem = i - part * g.natomFull * 128;
    const int a = rem >> 7, cc = (rem >> 4) & 7, kk = rem & 15;
    const int q = 8 * a + kk + cc - (kT2P - 1);
    __half v = __float2half_rn(0.f);
    if (q >= 0 && q < W) {
      const float t = tp[2 * q];
      const __half t1 = __float2half_rn(t);
      v = part == 0 ? t1 : __float2half_rn((t - __half2float(t1)) * kTcLoScale);
    }
    const size_t byteOff = (size_t)c * g.tapsFullBytes + (size_t)part * g.natomFull * 256 + (size_t)a * 256 + (size_t)cc * 32 +
                           (size_t)(((kk >> 3) ^ ((cc >> 2) & 1)) << 4) + (size_t)(kk & 7) * 2;
    *reinterpret_cast<__half *>(out + byteOff) = v;
  }
}

// centred query in Double for the exact re-evaluation (corr_fix.cuh), built on the device from the normalised Float query
// (same reason as the taps image: only a few KB cross PCIe at job creation): ac[c][i] = (double)a[c][i] + (-group mean)
__global__ void k_t2_query(const float *__restrict__ a, int numCh, int W, double negMeanT, double negMeanS, double *__restrict__ ac) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numCh * W) ac[i] = __dadd_rn((double)a[i], i < W ? negMeanT : negMeanS);
}

// tileFile[t] = file that holds frame min(8192 t, usedFrames - 1) (the file table is on the device already)
__global__ void k_t2_tile_files(const int64_t *__restrict__ fileStart, int numFiles, int64_t usedFrames, int64_t numTiles,
                                int32_t *__restrict__ tileFile) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t > numTiles) return;
  int64_t g = t * kT2Tile;
  if (g > usedFrames - 1) g = usedFrames - 1;
  if (g < 0) g = 0;
  int lo = 0, hi = numFiles;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (fileStart[mid] <= g) lo = mid; else hi = mid;
  }
  tileFile[t] = lo;
}

// ---------------------------------------------------------------------------------------------
// K0b: FP16 planes + per-frame sums of a frame range of the database (once per database / upload chunk)
// ---------------------------------------------------------------------------------------------
// planes[(2 c + part)][planeStrideBytes]: part 0 = fp16(x), part 1 = fp16((x - part0) * 2^11), pre-swizzled (t2_plane_byte);
// side arrays (tile transposed, t2_side_index): per frame b0 = loudness, s1 = sum over the spectral channels, s2 = sum of
// their squares (accumulated in FP64, rounded once), each stored as a Float -- K1 slides its window statistics in FP32,
// centred on the mean of the thread's first window (corr_tc2 epilogue); sideA = (b0, s1), sideB = s2;
// b16[(4 q + cb)][rowsTotal]: FP64 sums of (b0, b0^2, s1, s2)[q] of the SAME Float values over the aligned 16-frame
// block cb of each 64-frame row (exact: the window sums a thread starts from are the sums of what it slides with).
constexpr int kPlaneFrames = 2048;     // frames per block (256 threads x 8 frames)
__global__ void __launch_bounds__(256) k_db_planes(const float2 *__restrict__ data, int64_t rowStride, int numCh, int numPairs,
                                                   int64_t frameBegin, int64_t frameEnd, unsigned char *__restrict__ planes,
                                                   int64_t planeStrideBytes, uint2 *__restrict__ sideA,
                                                   uint32_t *__restrict__ sideB, double *__restrict__ b16, int64_t rowsTotal) {
  __shared__ uint32_t sh[3][kPlaneFrames + 32];    // +1 per 64 frames against bank conflicts of the transposed read-out
  const int64_t f0 = frameBegin + (int64_t)blockIdx.x * kPlaneFrames;   // frameBegin is a multiple of 2048
  const int64_t g0 = f0 + 8 * (int64_t)threadIdx.x;
  float b0[8];
  double s1[8], s2[8];
  for (int pr = 0; pr < numPairs; pr++) {
    const float4 *src = reinterpret_cast<const float4 *>(data + (int64_t)pr * rowStride + g0);
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = g0 + 2 * k < frameEnd ? __ldg(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    float x[8], y[8];
#pragma unroll
    for (int k = 0; k < 4; k++) { x[2 * k] = v[k].x; y[2 * k] = v[k].y; x[2 * k + 1] = v[k].z; y[2 * k + 1] = v[k].w; }
    if (pr == 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) { b0[k] = x[k]; s1[k] = (double)y[k]; s2[k] = (double)y[k] * (double)y[k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const double xd = (double)x[k], yd = (double)y[k];    // a channel beyond numCh is zero
        s1[k] += xd + yd; s2[k] = fma(xd, xd, fma(yd, yd, s2[k]));
      }
    }
    uint32_t xh[4], xl[4], yh[4], yl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const __half2 xa = __floats2half2_rn(x[2 * k], x[2 * k + 1]), ya = __floats2half2_rn(y[2 * k], y[2 * k + 1]);
      const float2 xf = __half22float2(xa), yf = __half22float2(ya);
      const __half2 xb = __floats2half2_rn((x[2 * k] - xf.x) * kTcLoScale, (x[2 * k + 1] - xf.y) * kTcLoScale);
      const __half2 yb = __floats2half2_rn((y[2 * k] - yf.x) * kTcLoScale, (y[2 * k + 1] - yf.y) * kTcLoScale);
      xh[k] = *reinterpret_cast<const uint32_t *>(&xa); xl[k] = *reinterpret_cast<const uint32_t *>(&xb);
      yh[k] = *reinterpret_cast<const uint32_t *>(&ya); yl[k] = *reinterpret_cast<const uint32_t *>(&yb);
    }
    const int64_t pos = t2_plane_byte(g0);     // 8 frames = one 16-byte chunk
    const int c0 = 2 * pr, c1 = 2 * pr + 1;
    *reinterpret_cast<uint4 *>(planes + (int64_t)(2 * c0) * planeStrideBytes + pos) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
    *reinterpret_cast<uint4 *>(planes + (int64_t)(2 * c0 + 1) * planeStrideBytes + pos) = make_uint4(xl[0], xl[1], xl[2], xl[3]);
    if (c1 < numCh) {
      *reinterpret_cast<uint4 *>(planes + (int64_t)(2 * c1) * planeStrideBytes + pos) = make_uint4(yh[0], yh[1], yh[2], yh[3]);
      *reinterpret_cast<uint4 *>(planes + (int64_t)(2 * c1 + 1) * planeStrideBytes + pos) = make_uint4(yl[0], yl[1], yl[2], yl[3]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int L = 8 * threadIdx.x + k, i = L + (L >> 6);
    sh[0][i] = __float_as_uint(b0[k]); sh[1][i] = __float_as_uint((float)s1[k]); sh[2][i] = __float_as_uint((float)s2[k]);
  }
  __syncthreads();
  // transposed write-out: the block's 2048 frames are 32 rows x 64 columns of one tile; a warp writes 32 consecutive rows
  // of one column (128 contiguous bytes)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = t2_side_index(f0);      // row of f0 within its tile, column 0
  for (int col = warp; col < 64; col += 8) {
    const int L = 64 * lane + col, i = L + (L >> 6);
    const int64_t o = base + ((int64_t)col << 7) + lane;
    sideA[o] = make_uint2(sh[0][i], sh[1][i]); sideB[o] = sh[2][i];
  }
  // FP64 sums of the 128 aligned 16-frame blocks: warp cb takes column block cb of the 32 rows, lane = row
  if (warp < 4) {
    const int L0 = 64 * lane + 16 * warp, i0 = L0 + (L0 >> 6);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const double x = (double)__uint_as_float(sh[0][i0 + e]), x2 = (double)__uint_as_float(sh[0][i0 + e + 1]);
      a0 += x; a1 += x * x; c0 += x2; c1 += x2 * x2;
      a2 += (double)__uint_as_float(sh[1][i0 + e]); c2 += (double)__uint_as_float(sh[1][i0 + e + 1]);
      a3 += (double)__uint_as_float(sh[2][i0 + e]); c3 += (double)__uint_as_float(sh[2][i0 + e + 1]);
    }
    double *dst = b16 + (int64_t)warp * rowsTotal + (f0 >> 6) + lane;
    dst[0] = a0 + c0; dst[4 * rowsTotal] = a1 + c1; dst[8 * rowsTotal] = a2 + c2; dst[12 * rowsTotal] = a3 + c3;
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float t2_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Double -> Float by truncation with integer instructions (F2F.F32.F64 and MUFU share the 16-lane conversion unit, and an
// offset needed eight of them): for finite values of Float's normal range, 2^-23 relative; smaller magnitudes give 0.
// Used for quantities whose last bit does not matter here (a variance on its way into rsqrt, a mean in a tiny correction).
__device__ __forceinline__ float t2_d2f(double d) {
  const uint32_t hi = (uint32_t)__double2hiint(d), lo = (uint32_t)__double2loint(d);
  const uint32_t a = hi & 0x7fffffffu;
  const uint32_t m = __funnelshift_l(lo, a - 0x38000000u, 3);
  return __uint_as_float(a >= 0x38100000u ? (m | (hi & 0x80000000u)) : 0u);
}

// constants of one evaluated offset, prepared on the host (kernel parameters = constant bank operands)
struct T2Eval {
  double nT, nS;            // cells of a window per group: W, (numCh - 1) W
  double invNTd, invNSd;    // 1 / n
  float nTf, nSf;
  float sqEpsT, sqEpsS;     // sqrt(eps / (1 - eps)) per group: a window whose variance is below eps of its mean square is
                            // re-evaluated exactly (corr_fix.cuh); eps = max(2e-3, 0.25 / n) -- the truncation of the
                            // split-FP16 products (2^-22 of the level per cell) averages out over the n cells of a window
                            // only, so windows of a few cells need a larger spread
  float gateNm;             // boost <= maxBoost  <=>  loudness sum of the window >= gateNm (calcBoost is monotone)
  float invNT2, invNS2;     // 1 / n^2
  float cT, cS;             // 1 / (n std_a) per group
  float kTn, kSn;           // rho / (n^2 std_a): correction for the rounded taps not summing to exactly zero
  float wT, wS;
  int useT, useS;
};

struct CorrT2Params {
  const unsigned char *planes;  // [numCh * 2][planeStrideBytes]
  int64_t planeStrideBytes;
  const uint2 *sideA;           // tile-transposed per-frame (loudness, spectral sum) as Double high words
  const uint32_t *sideB;        // ... spectral sum of squares
  const double *b16;            // [16][rowsTotal] FP64 sums of aligned 16-frame blocks
  int64_t rowsTotal;
  int64_t usedFrames;
  int numCh, W;
  const unsigned char *taps;    // k_t2_taps image
  T2Eval ev;
  const int64_t *fileStart;
  const int32_t *tileFile;      // [numTiles + 1] file that holds frame 8192 * tile (clamped to the last file)
  int numFiles, tailExtra;
  int64_t tileBegin, tileEnd;
  float *sim;
  unsigned long long *fileMax;
  uint32_t *fixList, *fixCount; // offsets whose windows are ill-conditioned (corr_fix.cuh)
  uint32_t fixCap;
  long long *prof;              // SGZ_CORR_TC_PROF: per CTA 24 cycle counters (k_corr_tc2<true>), or nullptr
  int smemMax;                  // the geometry (ring depths) is a function of (W, shared memory limit); computed per thread --
                                // passed as a parameter block it cost the epilogue more registers (ptxas)
  int ahead;                    // channels between the L2 prefetch of a signal stage and its bulk copy (0: no prefetch)
  int ring;                     // initial ring depths (signal << 4 | taps) of t2_geom
  int tapsFirst;                // developer knob: request the taps of a channel before its planes
  int narrow;                   // 1: band-limited MMAs (narrower N) in the first / last K steps of a window
  int l2hint;                   // bit 0: signal planes evict_first, bit 1: taps + block sums evict_last, bit 2: streaming curve stores
  int splitRelease;             // 1: the epilogue hands the two temporal accumulators back before it reads the spectral ones
  int dbg;                      // developer knob (SGZ_T2_DBG, profiling build only): 1 = no MMAs, 2 = no per-offset work, 4 = no curve stores, 8 = no per-frame loads, 16 = no window slides, 32 = no evaluation
};

// pull a range into L2 without a destination: the ring in shared memory then waits for L2, not for HBM, so the HBM latency
// is covered by bytes in flight that need no shared memory
__device__ __forceinline__ void t2_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool t2_test(uint64_t *bar, uint32_t parity) {   // non-blocking phase test
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// Window statistics of a thread's run of 16 offsets, in FP32 and CENTRED on the mean mu of the run's first window (taken
// from exact FP64 sums): c1 = sum (x - mu), c2 = sum (x - mu)^2 over the cells of the window.  n^2 var = n c2 - c1^2 then
// cancels only by 1 + (mean - mu)^2 / var (the mean moves by a few frames out of W during 15 slides), where the raw sums
// n sum x^2 - (sum x)^2 cancel by mean^2 / var; nmu = n mu, tmu = 2 mu.
struct T2Win {
  float c1T, c2T, c1S, c2S, nmuT, nmuS, tmuT, tmuS;
};

// one evaluated offset: centred window sums -> sim (FeatureCorrelationImpl.scala:73-78,198-210); branch free so that the
// compiler can interleave the offsets of a thread.  The boost itself is not needed here, only its gate (the few offsets
// that end up in a result get their boost from corr_boost, corr_fix.cuh).  A window is handed to the exact re-evaluation
// (sentinel NaN) when its variance is below eps of its mean square (the split-FP16 products truncate at the scale of the
// level) or when the centring has drifted (c1^2 above half of n c2: FP32 would cancel).
__device__ __forceinline__ float t2_eval(const T2Eval &E, const T2Win &w, float accT, float accS) {
  const float qnan = __uint_as_float(kFixSentinel);
  const float nmT = w.nmuT + w.c1T;                                     // n * window mean
  const float qT = w.c1T * w.c1T, vT = fmaf(E.nTf, w.c2T, -qT);         // n^2 var
  const float tT = nmT * E.sqEpsT;
  const float crT = fmaf(accT, E.cT, -(nmT * E.kTn));                   // (acc - mean_b * rho) / (W std_a)
  float temporal = crT * t2_rsqrt(vT * E.invNT2);
  temporal = vT > fmaxf(tT * tT, qT) ? temporal : qnan;                 // (false for NaN)
  temporal = E.useT ? temporal : 0.f;
  const float nmS = w.nmuS + w.c1S;
  const float qS = w.c1S * w.c1S, vS = fmaf(E.nSf, w.c2S, -qS);
  const float tS = nmS * E.sqEpsS;
  const float crS = fmaf(accS, E.cS, -(nmS * E.kSn));
  float spectral = crS * t2_rsqrt(vS * E.invNS2);
  spectral = vS > fmaxf(tS * tS, qS) ? spectral : qnan;
  spectral = E.useS ? spectral : 0.f;
  const float blend = __fadd_rn(__fmul_rn(temporal, E.wT), __fmul_rn(spectral, E.wS));
  // an ill-conditioned group makes the blend NaN whatever the other group is; it is marked for the exact re-evaluation
  return nmT >= E.gateNm ? (blend == blend ? blend : qnan) : 0.f;
}

// warp 0: producer (bulk copies), warp 1: MMA issuer, warps 2-3: idle (they complete the first warpgroup, which hands most
// of its registers to the others: setmaxnreg), warps 4..19: epilogue
// kMulti: long windows, several passes per tile (T2Geom::NP > 1)
// kFast: FILTER mode (one-pass windows only) -- the first-part product a1 t1 alone: one MMA of three, the first-part plane
// and taps alone (44 instead of 75 B/offset).  The sims are then accurate to a few 1e-4 instead of 1e-6; a punch-in search
// with numPerFile = 1 still returns the reference's matches bit for bit, because everything within the (wider) margin of
// the threshold is re-evaluated exactly afterwards (corr_refine.cuh).  Opt-in (SGZ_FAST=1).
template <bool kProf, bool kMulti, bool kFast = false>
__global__ void __launch_bounds__(kT2Threads, 1) k_corr_tc2(const CorrT2Params p) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  const T2Geom G = t2_geom(p.W, (size_t)p.smemMax, p.ring);
  const int NP = kMulti ? G.NP : 1;
  unsigned char *base = smemRaw + ((1024 - (smem_u32(smemRaw) & 1023)) & 1023);
  auto sigBuf = [&](int s, int part) { return base + (size_t)(2 * s + part) * G.planeStride; };
  unsigned char *tapsBase = base + (size_t)G.sigStages * 2 * G.planeStride;
  auto tapBuf = [&](int s) { return tapsBase + (size_t)s * G.tapsStride; };
  double *B16 = reinterpret_cast<double *>(tapsBase + (size_t)G.tapStages * G.tapsStride);   // [q][cb][sumPitch]
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(B16) + G.sumsBytes);
  uint64_t *sigFull = bars, *sigFree = bars + 8, *tapFull = bars + 16, *tapFree = bars + 20;
  uint64_t *accFull = bars + 24, *accEmpty = bars + 25, *sumsFull = bars + 26, *sumsFree = bars + 27;
  uint32_t *tmemSlot = reinterpret_cast<uint32_t *>(bars + 28);
#define accEmptyT (accEmpty + 4)      // accumulators 0 and 1 (temporal channel) drained; accEmpty: 2..7 (spectral channels)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 8; s++) { mbar_init(sigFull + s, 1); mbar_init(sigFree + s, 1); }
    for (int s = 0; s < 4; s++) { mbar_init(tapFull + s, 1); mbar_init(tapFree + s, 1); }
    mbar_init(accFull, 1);
    mbar_init(accEmpty, kT2EpiWarps);
    mbar_init(accEmptyT, kT2EpiWarps);
    mbar_init(sumsFull, 1);
    mbar_init(sumsFree, kT2EpiWarps);
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
  // 640 threads are launched with 96 registers each; the service warpgroup keeps 64, the epilogue warpgroups take 104 (the pool only holds what was released)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // =========================== producer ===========================
    if (lane == 0) {
      uint32_t it = 0, tileIt = 0;
      // the planes stream through once (1.7 GB per 1000 h / 10): marked evict_first they leave the per-frame arrays (touched
      // twice a tile apart), the taps and the block sums alone in L2
      const uint64_t polSig = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
      for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, tileIt++) {
        const unsigned char *src = p.planes + tile * (int64_t)(kT2Tile * 2);
        bool sumsPending = true;     // the block sums of this tile overwrite those of the previous one: once its windows are set up
        for (int j = 0; j < NP; j++) {
        const bool lastPass = j == NP - 1;
        const uint32_t shift = (uint32_t)j * (uint32_t)(G.KS * 32);   // pass j reads the signal 16 KS frames further on
        const uint32_t dOff = (j & 1) ? 512u : 0u;                    // ... which starts 4 j rows into the swizzle period
        for (int c = 0; c <= p.numCh; c++) {
          if (sumsPending && ((c == p.numCh && lastPass) || tileIt == 0 || t2_test(sumsFree, (tileIt - 1) & 1))) {
            if (tileIt > 0) tc_wait(sumsFree, (tileIt - 1) & 1);
            mbar_expect_tx(sumsFull, G.sumsBytes);
            for (int k = 0; k < 16; k++)
              bulk_g2s(B16 + k * G.sumPitch, p.b16 + (int64_t)k * p.rowsTotal + tile * kT2M, (uint32_t)G.sumPitch * 8u, sumsFull);
            sumsPending = false;
          }
          if (c == p.numCh) break;
          if (p.ahead > 0) {   // planes of the channel `ahead` steps ahead -> L2
            const int ca = c + p.ahead;
            const int64_t ta = tile + (int64_t)gridDim.x * (ca / p.numCh);
            if (ta < p.tileEnd) {
              const unsigned char *pa = p.planes + ta * (int64_t)(kT2Tile * 2) + (int64_t)(2 * (ca % p.numCh)) * p.planeStrideBytes;
              t2_prefetch_l2(pa, G.planeBytes);
              t2_prefetch_l2(pa + p.planeStrideBytes, G.planeBytes);
            }
            if (c == 0 && tile + gridDim.x < p.tileEnd)          // and the block sums of the next tile
              for (int k = 0; k < 16; k++)
                t2_prefetch_l2(p.b16 + (int64_t)k * p.rowsTotal + (tile + gridDim.x) * kT2M, (uint32_t)G.sumPitch * 8u);
          }
          auto issue_sig = [&]() {
            const uint32_t s = it % (uint32_t)G.sigStages, u = it / (uint32_t)G.sigStages;
            if (u > 0) tc_wait(sigFree + s, (u - 1) & 1);
            mbar_expect_tx(sigFull + s, (kFast ? 1u : 2u) * G.planeBytes);
            const unsigned char *s0 = src + (int64_t)(2 * c) * p.planeStrideBytes + shift, *s1 = s0 + p.planeStrideBytes;
            if ((p.l2hint & 1) && lastPass) {     // earlier passes: the planes of this tile come again
              bulk_g2s_hint(sigBuf(s, 0) + dOff, s0, G.planeBytes, sigFull + s, polSig);
              if (!kFast) bulk_g2s_hint(sigBuf(s, 1) + dOff, s1, G.planeBytes, sigFull + s, polSig);
            } else {
              bulk_g2s(sigBuf(s, 0) + dOff, s0, G.planeBytes, sigFull + s);
              if (!kFast) bulk_g2s(sigBuf(s, 1) + dOff, s1, G.planeBytes, sigFull + s);
            }
          };
          auto issue_taps = [&]() {
            const uint32_t t = it % (uint32_t)G.tapStages, v = it / (uint32_t)G.tapStages;
            if (v > 0) tc_wait(tapFree + t, (v - 1) & 1);
            const uint32_t tapBytes = kFast ? (uint32_t)G.natom * 256u : G.tapsBytes;     // filter mode: the first-part atoms
            mbar_expect_tx(tapFull + t, tapBytes);
            const unsigned char *tsrc = p.taps + (size_t)c * G.tapsFullBytes + (size_t)j * (size_t)(2 * G.KS * 256);
            if (!kMulti) {
              if (p.l2hint & 2) bulk_g2s_hint(tapBuf(t), tsrc, tapBytes, tapFull + t, polKeep);
              else bulk_g2s(tapBuf(t), tsrc, tapBytes, tapFull + t);
            } else {   // the atoms [32 j, 32 j + natom) of both parts
              bulk_g2s(tapBuf(t), tsrc, (uint32_t)G.natom * 256u, tapFull + t);
              bulk_g2s(tapBuf(t) + (size_t)G.natom * 256, tsrc + (size_t)G.natomFull * 256, (uint32_t)G.natom * 256u, tapFull + t);
            }
          };
          if (p.tapsFirst) { issue_taps(); issue_sig(); } else { issue_sig(); issue_taps(); }
          it++;
        }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // D = F32, A = B = F16, both K-major, M = 128, N = 64 (narrower at the edges of the taps band, see below)
    auto make_idesc = [](uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((uint32_t)(kT2M >> 4) << 24); };
    const uint32_t idesc = make_idesc(kT2P), id16 = make_idesc(16), id32 = make_idesc(32), id48 = make_idesc(48);
    uint32_t it = 0, tileIt = 0, passIt = 0;
    long long cAcc = 0, cSig = 0, cTap = 0, cIssue = 0, cTotal = 0, tA = 0;
    unsigned long long nsTotal = 0;
    if (kProf) { cTotal = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(nsTotal)); }
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, tileIt++) {
      for (int j = 0; j < NP; j++, passIt++) {
      const int nk = j == NP - 1 ? G.KSlast : G.KS;
      const uint32_t dOff = (j & 1) ? 512u : 0u;
      uint32_t started = 0;     // bit i: accumulator i holds a partial sum of this pass
      if (kProf) tA = clock64();
      // the epilogue has drained the temporal accumulators (0, 1) of the previous pass; the spectral ones (2..7) are
      // only written from channel 1 on, 3 KS MMAs later, so the read-out of a pass overlaps the first MMAs of the next
      if (passIt > 0) tc_wait<false>(accEmptyT, (passIt - 1) & 1);
      if (kProf) cAcc += clock64() - tA;
      for (int c = 0; c < p.numCh; c++, it++) {
        if (c == 1 && passIt > 0) {
          if (kProf) tA = clock64();
          tc_wait<false>(accEmpty, (passIt - 1) & 1);
          if (kProf) cAcc += clock64() - tA;
        }
        const uint32_t s = it % (uint32_t)G.sigStages, u = it / (uint32_t)G.sigStages;
        const uint32_t t = it % (uint32_t)G.tapStages, v = it / (uint32_t)G.tapStages;
        if (kProf) tA = clock64();
        tc_wait<false>(tapFull + t, v & 1);
        if (kProf) { cTap += clock64() - tA; tA = clock64(); }
        tc_wait<false>(sigFull + s, u & 1);
        if (kProf) { cSig += clock64() - tA; tA = clock64(); }
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t iMain = c == 0 ? 0u : 3u + (uint32_t)((c - 1) % kT2Mains), iCorr = c == 0 ? 1u : 2u;
        const uint32_t dMain = tmem + 64u * iMain, dCorr = tmem + 64u * iCorr;
        const uint32_t accMain = (started >> iMain) & 1u, accCorr = (started >> iCorr) & 1u;
        started |= (1u << iMain) | (1u << iCorr);
        const uint64_t aHi = tc_desc(smem_u32(sigBuf(s, 0)) + dOff, 1024, 2), aLo = tc_desc(smem_u32(sigBuf(s, 1)) + dOff, 1024, 2);
        const uint32_t tHiA = smem_u32(tapBuf(t));
        const uint64_t tHi = tc_desc(tHiA, 256, 6), tLo = tc_desc(tHiA + (uint32_t)G.natom * 256u, 256, 6);
        if (tc_elect()) {
          // one K step of 16: A start +32 B (2 descriptor units), taps start +2 atoms = 512 B (32 units); a1 t1 and
          // a1 t2 share the A operand through the collector.
          // The taps matrix is a band: in K step S only the columns c with 0 <= 16 S + kk + c - 63 < W hold taps, i.e.
          // c in [48 - 16 S, 62 + W - 16 S].  The first three and the last three K steps of a window therefore run as
          // narrower MMAs -- N = 16, 32, 48 on columns 48.., 32.., 16.. at the head, N = 48, 32, 16 on columns 0.. at the
          // tail: a fifth of the tensor work at W = 172 -- and the K step issued FIRST is a full one (k = 3), because it
          // may have to overwrite the accumulator.  (One-pass windows of at least 7 K steps; the issue loop is the
          // critical path of the kernel, so the six narrow steps are written out with constant geometry.)
          auto kstep = [&](uint32_t k, uint32_t c0, uint32_t id, uint32_t acc) {
            const uint64_t d1 = aHi + 2u * k, d2 = aLo + 2u * k, b1 = tHi + 32u * k + 2u * c0, b2 = tLo + 32u * k + 2u * c0;
            if (kFast) { tc_mma(dMain + c0, d1, b1, id, accMain | acc); return; }
            tc_mma_fill(dMain + c0, d1, b1, id, accMain | acc);
            tc_mma_lastuse(dCorr + c0, d1, b2, id, accCorr | acc);
            tc_mma(dCorr + c0, d2, b1, id, 1);
          };
          if (!kMulti && p.narrow && nk >= 7 && !(kProf && p.dbg & 1)) {
            uint64_t d1 = aHi + 6, d2 = aLo + 6, b1 = tHi + 96, b2 = tLo + 96;
            for (int k = 3; k <= nk - 4; k++, d1 += 2, d2 += 2, b1 += 32, b2 += 32) {
              if (kFast) { tc_mma(dMain, d1, b1, idesc, accMain | (uint32_t)(k > 3)); continue; }
              tc_mma_fill(dMain, d1, b1, idesc, accMain | (uint32_t)(k > 3));
              tc_mma_lastuse(dCorr, d1, b2, idesc, accCorr | (uint32_t)(k > 3));
              tc_mma(dCorr, d2, b1, idesc, 1);
            }
            kstep(0u, 48u, id16, 1u); kstep(1u, 32u, id32, 1u); kstep(2u, 16u, id48, 1u);
            kstep((uint32_t)nk - 3u, 0u, id48, 1u); kstep((uint32_t)nk - 2u, 0u, id32, 1u); kstep((uint32_t)nk - 1u, 0u, id16, 1u);
          } else {
            uint64_t d1 = aHi, d2 = aLo, b1 = tHi, b2 = tLo;
            for (int k = 0; k < nk; k++, d1 += 2, d2 += 2, b1 += 32, b2 += 32) {
              if ((kProf && p.dbg & 1)) break;
              if (kFast) { tc_mma(dMain, d1, b1, idesc, accMain | (uint32_t)(k > 0)); continue; }
              tc_mma_fill(dMain, d1, b1, idesc, accMain | (uint32_t)(k > 0));
              tc_mma_lastuse(dCorr, d1, b2, idesc, accCorr | (uint32_t)(k > 0));
              tc_mma(dCorr, d2, b1, idesc, 1);
            }
          }
          tc_commit(sigFree + s);
          tc_commit(tapFree + t);
          if (c == p.numCh - 1) tc_commit(accFull);
        }
        __syncwarp();
        if (kProf) cIssue += clock64() - tA;
      }
      }
    }
    if (kProf && p.prof && lane == 0) {
      long long *o = p.prof + 24 * blockIdx.x;
      unsigned long long nsEnd;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(nsEnd));
      o[0] = clock64() - cTotal; o[1] = cAcc; o[2] = cSig; o[3] = cTap; o[4] = cIssue; o[5] = tileIt; o[6] = (long long)(nsEnd - nsTotal);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // =========================== epilogue ===========================
    // 16 warps: TMEM lane quarter q = warp % 4 (rows r = 32 q + lane), column group j = (warp - 4) / 4.  A thread owns the
    // 16 offsets t0 + 64 r + jb .. + 15 (accumulator columns 63 - offset), jb = 16 (3 - j).
    const int ew = warp - 4, quarter = warp & 3, jg = ew >> 2, et = ew * 32 + lane;
    const int jb = 16 * (3 - jg);
    const int W = p.W;
    const T2Eval &E = p.ev;
    const float qnan = __int_as_float(0x7fc00000);
    const int r = quarter * 32 + lane;
    // window of the thread's first offset = nbk aligned 16-frame blocks, plus (sgn > 0) or minus (sgn < 0) nsg single frames
    const int nb = W >> 4, left = W & 15;
    const int nbk = left > 8 ? nb + 1 : nb, nsg = left > 8 ? 16 - left : left, sg0 = left > 8 ? W : nb << 4;
    const double sgn = left > 8 ? -1.0 : 1.0;
    uint32_t tileIt = 0, passIt = 0;
    long long eAcc = 0, eLd = 0, eSt = 0, eInit = 0, eMain = 0, tE = 0;
    for (int64_t tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x, tileIt++) {
      const int64_t t0 = tile * kT2Tile;
      const int64_t g0 = t0 + 64 * (int64_t)r + jb;
      {   // pull the per-frame arrays of the NEXT tile into L2: this tile's slides then wait for L2, not for HBM
        const int64_t nt = tile + gridDim.x;
        if (nt < p.tileEnd) {
          for (int l = et; l < 3 * 256; l += kT2EpiWarps * 32) {     // 512 lines of sideA, 256 of sideB
            const void *a = l < 512 ? (const void *)(p.sideA + nt * kT2Tile + (int64_t)l * 16)
                                    : (const void *)(p.sideB + nt * kT2Tile + (int64_t)(l - 512) * 32);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
      }
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
      if (kProf) tE = clock64();

      // ---- window statistics of the first offset (needs no accumulator) ----
      D4 win = {0, 0, 0, 0};
      {
        // single frames first: their loads are in flight while the block sums arrive
        D4 w2 = {0, 0, 0, 0};
        for (int e = 0; e < nsg; e++) {
          const int w = jb + sg0 + e, rr = r + (w >> 6);
          const int64_t o = (tile + (rr >> 7)) * (int64_t)kT2Tile + ((int64_t)(w & 63) << 7) + (rr & 127);
          const uint2 a = __ldg(p.sideA + o);
          const double x = (double)__uint_as_float(a.x);
          w2.t1 += x; w2.t2 += x * x; w2.s1 += (double)__uint_as_float(a.y); w2.s2 += (double)__uint_as_float(__ldg(p.sideB + o));
        }
        tc_wait(sumsFull, tileIt & 1);
        if (kProf) { eSt += clock64() - tE; tE = clock64(); }
        D4 w3 = {0, 0, 0, 0};
        int b = jb >> 4;                                      // block index relative to row r: block b -> row r + b / 4, cb b % 4
        int k = 0;
        for (; k + 1 < nbk; k += 2, b += 2) {
          const double *o = B16 + (b & 3) * G.sumPitch + r + (b >> 2);
          const double *o2 = B16 + ((b + 1) & 3) * G.sumPitch + r + ((b + 1) >> 2);
          win.t1 += o[0]; win.t2 += o[4 * G.sumPitch]; win.s1 += o[8 * G.sumPitch]; win.s2 += o[12 * G.sumPitch];
          w3.t1 += o2[0]; w3.t2 += o2[4 * G.sumPitch]; w3.s1 += o2[8 * G.sumPitch]; w3.s2 += o2[12 * G.sumPitch];
        }
        if (k < nbk) {
          const double *o = B16 + (b & 3) * G.sumPitch + r + (b >> 2);
          win.t1 += o[0]; win.t2 += o[4 * G.sumPitch]; win.s1 += o[8 * G.sumPitch]; win.s2 += o[12 * G.sumPitch];
        }
        // the sums are in registers: the producer may fetch the next tile's block sums
        __syncwarp();
        if (lane == 0) mbar_arrive(sumsFree);
        win.t1 = (win.t1 + w3.t1) + sgn * w2.t1; win.t2 = (win.t2 + w3.t2) + sgn * w2.t2;
        win.s1 = (win.s1 + w3.s1) + sgn * w2.s1; win.s2 = (win.s2 + w3.s2) + sgn * w2.s2;
      }
      // centre on the mean of this window: mu as a Float, the centred sums from the exact ones in FP64
      T2Win cw;
      {
        const float muT = (float)(win.t1 * E.invNTd), muS = (float)(win.s1 * E.invNSd);
        const double mT = (double)muT, mS = (double)muS;
        cw.c1T = (float)fma(-E.nT, mT, win.t1);
        cw.c2T = (float)fma(mT, fma(E.nT, mT, -2.0 * win.t1), win.t2);
        cw.c1S = (float)fma(-E.nS, mS, win.s1);
        cw.c2S = (float)fma(mS, fma(E.nS, mS, -2.0 * win.s1), win.s2);
        cw.nmuT = E.nTf * muT; cw.nmuS = E.nSf * muS;
        cw.tmuT = 2.f * muT; cw.tmuS = 2.f * muS;
      }
      if (kProf) { eInit += clock64() - tE; tE = clock64(); }

      // ---- accumulators -> registers (long windows: the sum over the passes) ----
      const uint32_t laneAddr = tmem + ((uint32_t)(quarter * 32) << 16) + 16u * (uint32_t)jg;
      float accT[16], accS[16];
      for (int j = 0; j < NP; j++, passIt++) {
        tc_wait(accFull, passIt & 1);
        if (kProf) { eAcc += clock64() - tE; tE = clock64(); }
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t u[16], w[16];
        tc_ld16_nowait(laneAddr + 0 * 64, u);
        if (!kFast) tc_ld16_nowait(laneAddr + 1 * 64, w);     // (filter mode: the correction accumulators are never written)
        tc_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float v = kFast ? __uint_as_float(u[i]) : fmaf(__uint_as_float(w[i]), 1.0f / kTcLoScale, __uint_as_float(u[i]));
          accT[i] = j == 0 ? v : accT[i] + v;
        }
        if (p.splitRelease) {
          asm volatile("tcgen05.fence::before_thread_sync;");
          __syncwarp();
          if (lane == 0) mbar_arrive(accEmptyT);
        }
        if (!kFast) tc_ld16_nowait(laneAddr + 2 * 64, u);
        tc_ld16_nowait(laneAddr + 3 * 64, w);
        tc_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float v = kFast ? __uint_as_float(w[i]) : fmaf(__uint_as_float(u[i]), 1.0f / kTcLoScale, __uint_as_float(w[i]));
          accS[i] = j == 0 ? v : accS[i] + v;
        }
        const int nMain = min(p.numCh - 1, kT2Mains);
        for (int m = 1; m < nMain; m += 2) {
          tc_ld16_nowait(laneAddr + (uint32_t)(3 + m) * 64, u);
          if (m + 1 < nMain) tc_ld16_nowait(laneAddr + (uint32_t)(4 + m) * 64, w);
          tc_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i++) accS[i] += __uint_as_float(u[i]) + (m + 1 < nMain ? __uint_as_float(w[i]) : 0.f);
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) { mbar_arrive(accEmpty); if (!p.splitRelease) mbar_arrive(accEmptyT); }
        if (kProf) tE = clock64();
      }
      if (kProf) { eLd += clock64() - tE; tE = clock64(); }

      // ---- 16 offsets: evaluate, slide the window by one frame ----
      unsigned long long best = 0ull;
      const bool plain = g0 + 16 <= fStart + ((fEnd - fStart) - p.tailExtra - W + 1) && g0 + 16 <= p.usedFrames;
      const uint32_t tl0 = (uint32_t)(g0 - fStart);
      float bestS = -INFINITY;                                 // plain runs: largest sim and its first position
      int bestJ = -1;
      // leaving frames 64 r + jb + e sit at + 128 e from the first one; entering frames 64 r + jb + W + e likewise until
      // the column wraps into the next row (once in 16 frames at most), which may be the first row of the next tile
      const uint32_t oOld = (uint32_t)(t0 + (jb << 7) + r);    // element indices fit 32 bits (sgz_corr_scan checks)
      const int wN = jb + W, colN = wN & 63, rowN = r + (wN >> 6);
      const uint32_t oNew = (uint32_t)((tile + (rowN >> 7)) * (int64_t)kT2Tile + (colN << 7) + (rowN & 127));
      const int eWrap = 64 - colN;
      const uint32_t dWrap = (uint32_t)(((rowN & 127) == 127 ? kT2Tile - 127 : 1) - (64 << 7));
      // the frames that leave / enter the window during four slides, loaded one block of four AHEAD: the loads of block
      // b + 1 are in flight while block b is evaluated (the latency of these loads was the largest stall of the epilogue)
      uint2 oaN[4], naN[4];
      uint32_t ob2N[4], nb2N[4];
      auto load_block = [&](int blk) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int jj = 4 * blk + e;
          if ((kProf && p.dbg & 8)) { oaN[e] = make_uint2(0x3f000000u + jj, 0x40800000u); naN[e] = oaN[e]; ob2N[e] = nb2N[e] = 0x40000000u; }
          else if (jj < 15) {
            oaN[e] = __ldg(p.sideA + (oOld + (uint32_t)(jj << 7))); ob2N[e] = __ldg(p.sideB + (oOld + (uint32_t)(jj << 7)));
            const uint32_t d = oNew + (uint32_t)(jj << 7) + (jj >= eWrap ? dWrap : 0u);
            naN[e] = __ldg(p.sideA + d); nb2N[e] = __ldg(p.sideB + d);
          }
        }
      };
      if (!((kProf && p.dbg & 2))) load_block(0);
#pragma unroll
      for (int blk = 0; blk < 4; blk++) {
        if ((kProf && p.dbg & 2)) break;
        uint2 oa[4], na[4];
        uint32_t ob2[4], nb2[4];
#pragma unroll
        for (int e = 0; e < 4; e++) { oa[e] = oaN[e]; na[e] = naN[e]; ob2[e] = ob2N[e]; nb2[e] = nb2N[e]; }
        if (blk < 3) load_block(blk + 1);
        float simv[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int jj = 4 * blk + e;           // offset inside the thread's run; accumulator column 15 - jj of its group
          simv[e] = ((kProf && p.dbg & 32)) ? accT[15 - jj] + accS[15 - jj] + cw.c1T : t2_eval(E, cw, accT[15 - jj], accS[15 - jj]);
          if (jj < 15 && !((kProf && p.dbg & 16))) {
            // (x - mu)^2 entering minus leaving = d (bn + bo - 2 mu); per frame sum_c (x - mu)^2 = s2 - 2 mu s1 + C mu^2
            const float bo = __uint_as_float(oa[e].x), bn = __uint_as_float(na[e].x), d = bn - bo;
            cw.c1T += d;
            cw.c2T = fmaf(d, (bn + bo) - cw.tmuT, cw.c2T);
            const float d1 = __uint_as_float(na[e].y) - __uint_as_float(oa[e].y);
            cw.c1S += d1;
            cw.c2S += fmaf(-cw.tmuS, d1, __uint_as_float(nb2[e]) - __uint_as_float(ob2[e]));
          }
        }
        if (plain) {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (simv[e] > bestS) { bestS = simv[e]; bestJ = 4 * blk + e; }      // NaN never wins, ties keep the first
            if (simv[e] != simv[e]) {                                           // rare: ill-conditioned window
              const uint32_t slot = atomicAdd(p.fixCount, 1u);
              if (slot < p.fixCap) p.fixList[slot] = (uint32_t)(g0 + 4 * blk + e);
            }
          }
        } else {
          for (int e = 0; e < 4; e++) {
            const int64_t g = g0 + 4 * blk + e;
            while (g >= fEnd && f + 1 < p.numFiles) {
              if (best != 0ull && p.fileMax) atomicMax(p.fileMax + f, best);
              best = 0ull;
              f++;
              fStart = fEnd;
              fEnd = p.fileStart[f + 1];
            }
            const int64_t tl = g - fStart;
            float sv = qnan;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) if (kk == e) sv = simv[kk];
            if (!(g < p.usedFrames && tl < (fEnd - fStart) - p.tailExtra - W + 1)) sv = qnan;
            else if (sv == sv) {
              const unsigned long long key = ((unsigned long long)float_order_key(sv) << 32) |
                                             (unsigned long long)(0xffffffffu - (uint32_t)tl);
              if (key > best) best = key;
            } else {
              const uint32_t slot = atomicAdd(p.fixCount, 1u);
              if (slot < p.fixCap) p.fixList[slot] = (uint32_t)g;
            }
#pragma unroll
            for (int kk = 0; kk < 4; kk++) if (kk == e) simv[kk] = sv;
          }
        }
        if (kProf && (p.dbg & 4)) {     // no stores: keep the values alive through a branch that is never taken
          if (simv[0] + simv[1] + simv[2] + simv[3] == 123.456f) p.sim[0] = 1.f;
        } else {
          if (p.l2hint & 4) __stcs(reinterpret_cast<float4 *>(p.sim + g0 + 4 * blk), make_float4(simv[0], simv[1], simv[2], simv[3]));
          else *reinterpret_cast<float4 *>(p.sim + g0 + 4 * blk) = make_float4(simv[0], simv[1], simv[2], simv[3]);
        }
      }
      if (bestJ >= 0)
        best = ((unsigned long long)float_order_key(bestS) << 32) | (unsigned long long)(0xffffffffu - (tl0 + (uint32_t)bestJ));
      // one atomic per WARP where its 32 runs lie in one file (almost always): 512 atomics per tile on the same 8 bytes --
      // and the neighbouring tiles on the same file -- serialise in L2 (0.28 ms of a 0.94 ms scan, tools/t2_ablate.py)
      if (p.fileMax) {
        const int f0 = __shfl_sync(0xffffffffu, f, 0);
        if (__all_sync(0xffffffffu, f == f0)) {
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
            best = o > best ? o : best;
          }
          if (lane == 0 && best != 0ull) atomicMax(p.fileMax + f0, best);
        } else if (best != 0ull) {
          atomicMax(p.fileMax + f, best);
        }
      }
      if (kProf) eMain += clock64() - tE;
    }
    if (kProf && p.prof && ew == 0 && lane == 0) {
      long long *o = p.prof + 24 * blockIdx.x + 8;
      o[0] = eSt; o[1] = eInit; o[2] = eAcc; o[3] = eLd; o[4] = eMain;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

#undef accEmptyT
}  // namespace sgz
