// cross.cuh -- K5: CrossSimilarity (CrossSimilarityImpl.scala:32-187), SURVEY.md section 8(f) rank 1.
//
// The reference slides the SHORTER file (the template, read completely, length L) over the longer one and
// writes one sim per step into a 1-channel audio file.  Its ring buffer is peculiar and the kernel replays it
// as it behaves, not as it was probably meant:
//   * the buffer has 8192 frames; the first read takes c0 = min(len2, 8192) frames at once, every later read ONE
//     frame, stored at readOff which wraps modulo L (:140-141,165)  ->  ring position p < L holds, at output k,
//     the newest frame c0 + m (m <= k-1) with (c0 + m) % L == p, or still the initial frame p;
//   * MathUtil.correlate wraps its read index modulo the BUFFER length (MathUtil.scala:189), so logical frame i
//     of output k is buffer position (i + k % L) % 8192 -- positions >= L are the untouched frames of the first
//     read (or 0.0f where nothing was read);
//   * mean / std-dev / loudness average of the window always come from positions [0, L) (:181-185);
//   * 1 + len2 - c0 output values.
// One thread per output value replays the Double arithmetic operation by operation (stat two-pass in physical
// order, correlate in logical order, no FMA contraction), like K3 -> the curve is bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace sgz {

constexpr int kCrossBuf = 8192;

struct CrossParams {
  const float *x;        // normalised planar [numCh][stride] frames of the LONGER span (afIn2), index 0 = span start
  int64_t stride;
  const double *a;       // centred template [numCh][L]: (double)a[c][i] + (-mean of its group), the first factor of
                         // MathUtil.correlate's product, formed once on the host (same IEEE addition)
  int numCh;
  int L;                 // len1i
  int c0;                // frames of the first read = min(len2, 8192)
  int64_t nOut;
  double meanT, stdT, meanS, stdS;   // MathUtil.stat of the template groups
  double lnAvgIn;
  float weight, maxBoost;
  float *sim;            // [nOut]
};

// ring state of output k: which source frame sits at buffer position p (< L)
struct CrossRing {
  int L, r, c0;
  long long a;   // (k-1) / L
  int b;         // (k-1) % L
  bool any;      // k >= 1
  __device__ CrossRing(const CrossParams &p, long long k) : L(p.L), r(p.c0 % p.L), c0(p.c0) {
    any = k >= 1;
    a = any ? (k - 1) / L : 0;
    b = any ? (int)((k - 1) % L) : 0;
  }
  // with q = (pos - r) mod L the first m >= 0 written to position pos, that position holds frame
  //   c0 + q + a L        if k >= 1 and q <= b,
  //   c0 + q + (a - 1) L  if k >= 1, q > b and a >= 1,
  //   pos                 otherwise (still the frame of the first read)      -> cross_ring_runs
};

// Calls f(frame0, n) for the maximal runs of consecutive ring positions pos0, pos0 + 1, ... (count of them, all < L)
// that hold consecutive source frames frame0, frame0 + 1, ...: at most four runs per sweep (the write pointer and the
// wrap of q), so the sweeps below are plain linear loops without per-element index arithmetic.
template <typename F>
__device__ __forceinline__ void cross_ring_runs(const CrossRing &ring, int pos0, int count, F f) {
  int pos = pos0, left = count;
  int q = pos0 - ring.r;                       // q = (pos - r) mod L
  if (q < 0) q += ring.L;
  while (left > 0) {
    int n = min(left, ring.L - q);
    long long frame;
    if (ring.any && q <= ring.b) {
      n = min(n, ring.b + 1 - q);
      frame = (long long)ring.c0 + q + ring.a * ring.L;
    } else if (ring.any && ring.a >= 1) {
      frame = (long long)ring.c0 + q + (ring.a - 1) * ring.L;
    } else {
      frame = pos;                             // still the frames of the first read
    }
    f(frame, n);
    pos += n;
    left -= n;
    q += n;
    if (q == ring.L) q = 0;
  }
}

// FeatureMatrix a (channels [chanOff, chanOff+numChannels)) against the ring at output k
__device__ float cross_correlate(const CrossParams &p, const CrossRing &ring, int o, int chanOff, int numChannels,
                                 double aStd) {
  const int L = p.L;
  const int matSize = numChannels * L;
  // MathUtil.stat(b, 0, L, chanOff, numChannels): physical positions 0..L-1
  double sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const float *xc = p.x + (int64_t)(ch + chanOff) * p.stride;
    cross_ring_runs(ring, 0, L, [&](long long frame, int n) {
      const float *src = xc + frame;
      for (int j = 0; j < n; j++) sum = __dadd_rn(sum, (double)src[j]);
    });
  }
  const double bMean = __ddiv_rn(sum, (double)matSize);
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const float *xc = p.x + (int64_t)(ch + chanOff) * p.stride;
    cross_ring_runs(ring, 0, L, [&](long long frame, int n) {
      const float *src = xc + frame;
      for (int j = 0; j < n; j++) {
        const double d = __dsub_rn((double)src[j], bMean);
        sum = __dadd_rn(sum, __dmul_rn(d, d));
      }
    });
  }
  const double bStd = __dsqrt_rn(__ddiv_rn(sum, (double)matSize));
  // MathUtil.correlate: b index (i + o) % 8192; positions >= L hold the frames of the first read (or fresh zeros)
  const double bAdd = -bMean;
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const double *ca = p.a + (int64_t)(ch + chanOff) * L;
    const float *xc = p.x + (int64_t)(ch + chanOff) * p.stride;
    int i = 0, pos = o;
    while (i < L) {
      if (pos < L) {
        const int n = min(L - i, L - pos);
        cross_ring_runs(ring, pos, n, [&](long long frame, int m) {
          const float *src = xc + frame;
          for (int j = 0; j < m; j++, i++)
            sum = __dadd_rn(sum, __dmul_rn(ca[i], __dadd_rn((double)src[j], bAdd)));
        });
        pos += n;
      } else {
        const int n = min(L - i, kCrossBuf - pos);
        for (int j = 0; j < n; j++, i++) {
          const float bv = pos + j < p.c0 ? xc[pos + j] : 0.0f;   // beyond the first read: fresh array zeros
          sum = __dadd_rn(sum, __dmul_rn(ca[i], __dadd_rn((double)bv, bAdd)));
        }
        pos += n;
      }
      if (pos == kCrossBuf) pos = 0;           // wraps only when L > 4096
    }
  }
  return (float)__ddiv_rn(sum, __dmul_rn(__dmul_rn(aStd, bStd), (double)matSize));
}

__global__ void k_cross(const CrossParams p) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= p.nOut) return;
  const CrossRing ring(p, k);
  const int L = p.L;
  const int o = (int)(k % L);
  // calcBoost (:64-67): MathUtil.avg of the loudness channel over positions [0, L)
  double sum = 0.0;
  cross_ring_runs(ring, 0, L, [&](long long frame, int n) {
    const float *src = p.x + frame;
    for (int j = 0; j < n; j++) sum = __dadd_rn(sum, (double)src[j]);
  });
  const float avgB = (float)__ddiv_rn(sum, (double)L);
  const float boost = (float)exp(__ddiv_rn(__dsub_rn(p.lnAvgIn, log((double)avgB)), 0.6));
  float sim = 0.0f;
  if (boost <= p.maxBoost) {   // false for NaN, like the JVM
    const float temporal = p.weight > 0.f ? cross_correlate(p, ring, o, 0, 1, p.stdT) : 0.f;
    const float spectral = p.weight < 1.f ? cross_correlate(p, ring, o, 1, p.numCh - 1, p.stdS) : 0.f;
    sim = __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
  }
  p.sim[k] = sim;
}

}  // namespace sgz
