// selfsim_fast.cuh -- K4 (fast path): SelfSimilarity as a tiled Gram matrix.
//
// Closed form of MathUtil.correlateHalf on the buffer [window i | window j]:
//     cell_g(i,j) = (G - T) / ((Q_i + Q_j)/2 - T),   T = (S_i + S_j)^2 / (4 N),  N = C_g * H
// with G = <window i, window j> over the group's channels, S / Q = sum / sum of squares of a window.
// The expression is invariant under a common shift of all values, so the data are centred by one
// constant per group (the file mean) before the Gram: that removes the G ~ T cancellation and lets G
// be accumulated in FP32; S, Q, T and the denominator are FP64.
//
// G is a GEMM  A * B^T  whose rows are the (decimated) windows themselves:
//     A[a][k = c*H + h] = x[c][decim*a + h]     -- a Hankel matrix, never materialised: the tile
// loader gathers it from the planar feature rows.  128 x 128 cell tiles of the upper triangle,
// 256 threads, 8 x 8 cells per thread as 8 rows x 4 FFMA2 column pairs (same register-operand FFMA2
// pattern that reaches the full 73 TFLOP/s in tools/peaks_probe.py), k-chunks of 16 double buffered in
// shared memory.  Temporal group first (K = H), its coefficients parked in registers, then the
// spectral group (K = (C-1) H); blend, colour map and both mirrored pixel stores are fused.
#pragma once
#include "common.cuh"
#include "selfsim.cuh"

namespace sgz {

constexpr int kGT = 128;   // cells per tile side
constexpr int kGK = 16;    // k-chunk
constexpr int kGP = kGT + 4;  // smem row pitch (floats)

struct SelfFastParams {
  SelfParams base;         // x1/x2 planar normalised rows, geometry, colours, rgb
  float shiftT, shiftS;    // centring constants (file-1 group means)
  const float2 *ws1;       // [2][imgExt] (S, Q) of window a of file 1: group 0 = temporal, 1 = spectral
  const float2 *ws2;      // same for file 2 (== ws1 for plain self similarity)
  int cross;               // file 2 differs from file 1
};

// mean of channel 0 and of channels 1.. over the frames [0, n)
__global__ void k_self_means(const float *__restrict__ x, int64_t stride, int64_t n, int numCh, double *out) {
  double sT = 0, sS = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    sT += (double)x[i];
    for (int c = 1; c < numCh; c++) sS += (double)x[(int64_t)c * stride + i];
  }
  for (int d = 16; d > 0; d >>= 1) {
    sT += __shfl_xor_sync(0xffffffffu, sT, d);
    sS += __shfl_xor_sync(0xffffffffu, sS, d);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, sT); atomicAdd(out + 1, sS); }
}

// (S, Q) of every decimated window, centred data, FP64
__global__ void k_self_wsums(const float *__restrict__ x, int64_t stride, int numCh, int H, int decim, int imgExt,
                             float shiftT, float shiftS, float2 *__restrict__ ws) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= imgExt) return;
  const int64_t f0 = (int64_t)decim * a;
  double s = 0, q = 0;
  for (int h = 0; h < H; h++) { const double v = (double)__fsub_rn(x[f0 + h], shiftT); s += v; q += v * v; }
  ws[a] = make_float2((float)s, (float)q);
  s = 0; q = 0;
  for (int c = 1; c < numCh; c++)
    for (int h = 0; h < H; h++) { const double v = (double)__fsub_rn(x[(int64_t)c * stride + f0 + h], shiftS); s += v; q += v * v; }
  ws[imgExt + a] = make_float2((float)s, (float)q);
}

// Gather of the Hankel operand, one 16 x 128 chunk per call: thread t fetches the 8 rows i = 8*(t & 15) .. +7 at
// ONE k = k0 + (t >> 4), i.e. one (channel, frame) pair -> the 8 addresses differ by `decim` floats, there is no
// wrap logic per element, rows past the image are clamped (their cells are never stored) instead of predicated,
// and the 8 values go to shared memory as two 16-byte stores.  The (channel, frame) cursor advances by 16 per chunk.
struct GramCursor {
  const float *ptr;   // &x[c][decim*a0 + h] of this thread's k in the current chunk
  int64_t wrap;       // stride - H : jump to the next channel row when h wraps
  int64_t rstep[8];   // decim * (row offset), clamped at the image edge
  int h, k, H, K;
  float shift;
};
__device__ __forceinline__ GramCursor gram_cursor(const float *__restrict__ x, int64_t stride, int decim, int t0, int imgExt,
                                                  int c0, int H, int K, float shift, int tid) {
  GramCursor c;
  c.H = H; c.K = K; c.shift = shift;
  c.k = tid >> 4;
  const int cc = c.k / H;           // once per group
  c.h = c.k - cc * H;
  const int i0 = 8 * (tid & 15);
#pragma unroll
  for (int j = 0; j < 8; j++) c.rstep[j] = (int64_t)decim * min(t0 + i0 + j, imgExt - 1);
  c.ptr = x + (int64_t)(c0 + cc) * stride + c.h;
  c.wrap = stride - H;
  return c;
}
__device__ __forceinline__ void gram_fetch(float (&r)[8], GramCursor &c) {
  const bool ok = c.k < c.K;
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = ok ? __fsub_rn(c.ptr[c.rstep[j]], c.shift) : 0.f;
  c.k += kGK; c.h += kGK; c.ptr += kGK;
  while (c.h >= c.H) { c.h -= c.H; c.ptr += c.wrap; }
}
__device__ __forceinline__ void gram_store(float *dst, const float (&r)[8], int tid) {
  float4 *d = reinterpret_cast<float4 *>(dst + (tid >> 4) * kGP + 8 * (tid & 15));
  d[0] = make_float4(r[0], r[1], r[2], r[3]);
  d[1] = make_float4(r[4], r[5], r[6], r[7]);
}

// acc[i][j] (j = column pair) += A-frag x B-frag over one group
__device__ __forceinline__ void gram_group(float2 (&acc)[8][4], const SelfFastParams &p, float *sA, float *sB, int ta,
                                           int tb, int c0, int nC, float shift, int tid, int ty, int tx) {
  const SelfParams &b = p.base;
  const int K = nC * b.H;
  const int nChunks = (K + kGK - 1) / kGK;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);
  GramCursor ca = gram_cursor(b.x1, b.stride1, b.decim, ta, b.imgExt, c0, b.H, K, shift, tid);
  GramCursor cb = gram_cursor(b.x2, b.stride2, b.decim, tb, b.imgExt, c0, b.H, K, shift, tid);
  float ra[8], rb[8];
  gram_fetch(ra, ca);
  gram_fetch(rb, cb);
  gram_store(sA, ra, tid);
  gram_store(sB, rb, tid);
  __syncthreads();
  for (int ch = 0; ch < nChunks; ch++) {
    const float *cA = sA + (ch & 1) * kGK * kGP, *cB = sB + (ch & 1) * kGK * kGP;
    const bool more = ch + 1 < nChunks;
    if (more) {   // global loads of the next chunk stay in flight during this chunk's FFMA2 loop
      gram_fetch(ra, ca);
      gram_fetch(rb, cb);
    }
#pragma unroll
    for (int kk = 0; kk < kGK; kk++) {
      // rows: ty*4 + {0..3} and 64 + ty*4 + {0..3};  columns: tx*4 + {0..3} and 64 + tx*4 + {0..3}
      const float4 a0 = *reinterpret_cast<const float4 *>(cA + kk * kGP + ty * 4);
      const float4 a1 = *reinterpret_cast<const float4 *>(cA + kk * kGP + 64 + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4 *>(cB + kk * kGP + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4 *>(cB + kk * kGP + 64 + tx * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                            make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(aa, bv[j], acc[i][j]);
      }
    }
    if (more) {
      gram_store(sA + ((ch + 1) & 1) * kGK * kGP, ra, tid);
      gram_store(sB + ((ch + 1) & 1) * kGK * kGP, rb, tid);
    }
    __syncthreads();
  }
}

// inv4N = 1 / (4 N).  After centring, T is small against G and Q, so FP32 is enough here (error ~2e-7);
// the window sums S, Q themselves are accumulated in FP64 (k_self_wsums) and rounded once.
__device__ __forceinline__ float gram_coeff(float G, float2 wa, float2 wb, float inv4N) {
  const float S = wa.x + wb.x;
  const float T = S * S * inv4N;
  const float den = fmaf(0.5f, wa.y + wb.y, -T);
  return __fdiv_rn(G - T, den);   // 0/0 -> NaN like the reference (constant windows)
}

// cell-list twin of the tile kernel (parity checks): same centred FP32 Gram + FP64 closed form, one thread per cell
__global__ void k_self_cells_fast(const SelfFastParams p) {
  const SelfParams &b = p.base;
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= b.nCells) return;
  const int64_t l0 = (int64_t)b.leftIdx[k] * b.decim, r0 = (int64_t)b.rightIdx[k] * b.decim;
  float corr[2] = {0.f, 0.f};
  for (int g = 0; g < 2; g++) {
    if ((g == 0 && !(b.weight > 0.f)) || (g == 1 && !(b.weight < 1.f))) continue;
    const int c0 = g == 0 ? 0 : 1, c1 = g == 0 ? 1 : b.numCh;
    const float sh = g == 0 ? p.shiftT : p.shiftS;
    float G = 0.f;
    double sa = 0, qa = 0, sb = 0, qb = 0;
    for (int c = c0; c < c1; c++)
      for (int h = 0; h < b.H; h++) {
        const float u = __fsub_rn(b.x1[(int64_t)c * b.stride1 + l0 + h], sh);
        const float v = __fsub_rn(b.x2[(int64_t)c * b.stride2 + r0 + h], sh);
        G = fmaf(u, v, G);
        sa += (double)u; qa += (double)u * (double)u; sb += (double)v; qb += (double)v * (double)v;
      }
    corr[g] = gram_coeff(G, make_float2((float)sa, (float)qa), make_float2((float)sb, (float)qb),
                         (float)(1.0 / (4.0 * (double)(c1 - c0) * (double)b.H)));
  }
  const float sim = __fadd_rn(__fmul_rn(corr[0], b.weight), __fmul_rn(corr[1], __fsub_rn(1.0f, b.weight)));
  if (b.simOut) b.simOut[k] = sim;
  if (b.rgbOut) b.rgbOut[k] = self_color(b, sim);
}

__global__ void __launch_bounds__(256, 1) k_self_gram(const SelfFastParams p, const int2 *__restrict__ tiles) {
  __shared__ __align__(16) float sA[2 * kGK * kGP];
  __shared__ __align__(16) float sB[2 * kGK * kGP];
  const SelfParams &b = p.base;
  const int ta = tiles[blockIdx.x].x, tb = tiles[blockIdx.x].y;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int ext = b.imgExt;
  float2 acc[8][4];
  float corrT[8][8];
  int rowsA[8], colsB[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    rowsA[i] = ta + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    colsB[i] = tb + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
  }
  const bool useT = b.weight > 0.f, useS = b.weight < 1.f;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) corrT[i][j] = 0.f;
  if (useT) {
    gram_group(acc, p, sA, sB, ta, tb, 0, 1, p.shiftT, tid, ty, tx);
    const float N = (float)(1.0 / (4.0 * (double)b.H));
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float2 wa = rowsA[i] < ext ? p.ws1[rowsA[i]] : make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float2 wb = colsB[j] < ext ? p.ws2[colsB[j]] : make_float2(0.f, 0.f);
        const float G = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
        corrT[i][j] = gram_coeff(G, wa, wb, N);
      }
    }
  }
  if (useS) gram_group(acc, p, sA, sB, ta, tb, 1, b.numCh - 1, p.shiftS, tid, ty, tx);
  const float NS = (float)(1.0 / (4.0 * (double)(b.numCh - 1) * (double)b.H));
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int a = rowsA[i];
    if (a >= ext || a < b.colBegin || a >= b.colEnd) continue;
    const float2 wa = p.ws1[ext + a];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int c = colsB[j];
      if (c >= ext || c < a) continue;      // upper triangle only (also in cross mode, like the reference)
      float spectral = 0.f;
      if (useS) {
        const float G = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
        spectral = gram_coeff(G, wa, p.ws2[ext + c], NS);
      }
      const float sim = __fadd_rn(__fmul_rn(corrT[i][j], b.weight), __fmul_rn(spectral, __fsub_rn(1.0f, b.weight)));
      const int32_t colr = self_color(b, sim);
      b.rgb[(int64_t)(ext - 1 - c) * ext + a] = colr;
      b.rgb[(int64_t)(ext - 1 - a) * ext + c] = colr;
    }
  }
}

}  // namespace sgz
