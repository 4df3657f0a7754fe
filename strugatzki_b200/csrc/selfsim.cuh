// selfsim.cuh -- K4: SelfSimilarity matrix (SelfSimilarityImpl.scala:31-180).
//
// Round-1 kernel: every image cell is ONE thread that replays the reference's per-cell Double
// arithmetic (two MathUtil.correlateHalf calls, MathUtil.scala:80-99, on a buffer whose left half
// is file 1 at leftOff and right half is file 2 at rightOff), with the windows of a 16 x 16 cell
// tile staged in shared memory.  B200's FP64 pipe makes this exact path cheap enough to be the
// first correct version; it yields pixel-identical images.  DESIGN.md describes the planned
// phase-decomposed Gram formulation (K = 14*decim, diagonal box filter) that replaces it.
#pragma once
#include "common.cuh"

namespace sgz {

struct SelfParams {
  const float *x1;      // normalised planar [numCh][stride1], index 0 = afStart
  const float *x2;
  int64_t stride1, stride2;
  int numCh, H, decim, imgExt;
  int colBegin, colEnd; // image columns x (= leftOff/decim) computed by this launch
  float weight;
  int colorInv;
  float colorWarp, colorScale;
  const int32_t *lut;
  int lutSize;
  int32_t *rgb;         // [imgExt][imgExt] or nullptr
  // cell-list mode (parity checks)
  const int32_t *leftIdx, *rightIdx;
  int64_t nCells;
  float *simOut;
  int32_t *rgbOut;
};

constexpr int kSelfTile = 16;

__device__ __forceinline__ int32_t d2i_java(double d) {
  if (d != d) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (-2147483647 - 1);
  return (int32_t)d;
}

// colorFun(pow(max(0, sim), colorWarp).toFloat * colorScale), SelfSimilarityImpl.scala:98-111,150
__device__ inline int32_t self_color(const SelfParams &p, float sim) {
  float m = (sim != sim) ? sim : fmaxf(0.0f, sim);   // math.max(0f, NaN) = NaN
  if (sim == 0.0f) m = 0.0f;                          // max(0f, -0f) = +0f
  float v = __fmul_rn((float)pow((double)m, (double)p.colorWarp), p.colorScale);
  float s = p.colorInv ? __fsub_rn(1.0f, v) : v;
  if (p.lut == nullptr) {
    float f255 = __fmul_rn(s, 255.0f);
    int32_t i = d2i_java(__dadd_rn((double)f255, 0.5));
    i = max(0, min(255, i));
    return (i << 16) | (i << 8) | i;
  }
  // PsychoOptical: the host supplies IntensityPalette as a LUT over [0,1] (parity unpinned)
  int32_t i = d2i_java(__dadd_rn((double)__fmul_rn(s, (float)(p.lutSize - 1)), 0.5));
  i = max(0, min(p.lutSize - 1, i));
  return p.lut[i] & 0x00ffffff;
}

// correlateHalf on the virtual buffer [left window | right window], frameOff = 0
template <typename LoadL, typename LoadR>
__device__ float self_correlate_half(int H, int chanOff, int numChannels, LoadL ldL, LoadR ldR) {
  const int matFull = 2 * H * numChannels;
  double sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const int c = ch + chanOff;
    for (int i = 0; i < H; i++) sum = __dadd_rn(sum, (double)ldL(c, i));
    for (int i = 0; i < H; i++) sum = __dadd_rn(sum, (double)ldR(c, i));
  }
  const double mean = __ddiv_rn(sum, (double)matFull);
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const int c = ch + chanOff;
    for (int i = 0; i < H; i++) { double d = __dsub_rn((double)ldL(c, i), mean); sum = __dadd_rn(sum, __dmul_rn(d, d)); }
    for (int i = 0; i < H; i++) { double d = __dsub_rn((double)ldR(c, i), mean); sum = __dadd_rn(sum, __dmul_rn(d, d)); }
  }
  const double stdDev = __dsqrt_rn(__ddiv_rn(sum, (double)matFull));
  const double add = -mean;
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const int c = ch + chanOff;
    for (int i = 0; i < H; i++) {
      double a = __dadd_rn((double)ldL(c, i), add), b = __dadd_rn((double)ldR(c, i), add);
      sum = __dadd_rn(sum, __dmul_rn(a, b));
    }
  }
  return (float)__ddiv_rn(sum, __dmul_rn(__dmul_rn(stdDev, stdDev), (double)(numChannels * H)));
}

template <typename LoadL, typename LoadR>
__device__ inline float self_cell_sim(const SelfParams &p, LoadL ldL, LoadR ldR) {
  const float temporal = p.weight > 0.f ? self_correlate_half(p.H, 0, 1, ldL, ldR) : 0.f;
  const float spectral = p.weight < 1.f ? self_correlate_half(p.H, 1, p.numCh - 1, ldL, ldR) : 0.f;
  return __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
}

// tile (bx, by) of the upper triangle: columns a (left) in [colBegin + 16*bx, +16), rows b (right)
__global__ void __launch_bounds__(kSelfTile *kSelfTile) k_self_tiles(const SelfParams p, const int2 *tiles) {
  extern __shared__ float sh[];
  const int a0 = tiles[blockIdx.x].x, b0 = tiles[blockIdx.x].y;
  const int span = p.decim * (kSelfTile - 1) + p.H;   // frames staged per side
  float *sL = sh;                                     // [numCh][span]
  float *sR = sh + (size_t)p.numCh * span;
  for (int i = threadIdx.x; i < p.numCh * span; i += blockDim.x) {
    int c = i / span, e = i - c * span;
    int64_t gl = (int64_t)p.decim * a0 + e, gr = (int64_t)p.decim * b0 + e;
    sL[i] = gl < p.stride1 ? p.x1[(int64_t)c * p.stride1 + gl] : 0.f;
    sR[i] = gr < p.stride2 ? p.x2[(int64_t)c * p.stride2 + gr] : 0.f;
  }
  __syncthreads();
  const int ta = threadIdx.x % kSelfTile, tb = threadIdx.x / kSelfTile;
  const int a = a0 + ta, b = b0 + tb;
  if (a < p.colBegin || a >= p.colEnd || a >= p.imgExt || b >= p.imgExt || b < a) return;
  const float *bl = sL + p.decim * ta, *br = sR + p.decim * tb;
  auto ldL = [&](int c, int i) { return bl[c * span + i]; };
  auto ldR = [&](int c, int i) { return br[c * span + i]; };
  const float sim = self_cell_sim(p, ldL, ldR);
  const int32_t colr = self_color(p, sim);
  const int64_t ext = p.imgExt;
  p.rgb[(ext - 1 - b) * ext + a] = colr;   // off1, :152
  p.rgb[(ext - 1 - a) * ext + b] = colr;   // off2, :153
}

__global__ void k_self_cells(const SelfParams p) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= p.nCells) return;
  const int64_t l0 = (int64_t)p.leftIdx[k] * p.decim, r0 = (int64_t)p.rightIdx[k] * p.decim;
  auto ldL = [&](int c, int i) { return p.x1[(int64_t)c * p.stride1 + l0 + i]; };
  auto ldR = [&](int c, int i) { return p.x2[(int64_t)c * p.stride2 + r0 + i]; };
  const float sim = self_cell_sim(p, ldL, ldR);
  if (p.simOut) p.simOut[k] = sim;
  if (p.rgbOut) p.rgbOut[k] = self_color(p, sim);
}

}  // namespace sgz
