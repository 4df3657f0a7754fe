// segm.cuh -- K3: FeatureSegmentation (FeatureSegmentationImpl.scala:31-142).
//
// north_star asks for BIT-IDENTICAL break frames.  B200 has a full-rate FP64 pipe
// (64 DFMA lanes / SM / clk), and the whole job is 51 595 offsets x 3 passes x 86 x 14 cells
// = 1.9e8 FP64 operations, i.e. microseconds of pipe time -- so instead of a fast FP32 curve plus a
// tie analysis, the curve kernel replays the reference's Double arithmetic operation by operation
// (MathUtil.stat two-pass in PHYSICAL ring-buffer order, MathUtil.correlateHalf in logical order,
// no FMA contraction: __dadd_rn/__dmul_rn), which makes every sim bit-identical to the JVM's and
// the break picking a pure replay of addBreak (:68-83).
#pragma once
#include "common.cuh"

namespace sgz {

struct SegmParams {
  const float *x;       // normalised planar [numCh][stride] frames of the analysed span (index 0 = afStart)
  int64_t stride;
  int afLen;            // frames available in the span
  int numCh;
  int H;                // halfWinLen
  int nOff;             // number of offsets
  float weight;
  float *curve;         // [nOff]
};

__device__ __forceinline__ float segm_load(const SegmParams &p, int c, int e) {
  // frames the reference never read stay 0.0f in its freshly allocated ring buffer (afLen < 2H case)
  return e < p.afLen ? p.x[(int64_t)c * p.stride + e] : 0.0f;
}

// MathUtil.correlateHalf(numChannels, H, ring, frameOff = t % 2H, chanOff) for window start t
__device__ float segm_correlate_half(const SegmParams &p, int t, int chanOff, int numChannels) {
  const int H = p.H, win = 2 * H;
  const int phase = t % win;
  const int matFull = win * numChannels;
  // stat pass 1: sum over PHYSICAL ring index phi = 0..2H-1; logical frame = t + ((phi - t) mod 2H)
  double sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const int c = ch + chanOff;
    for (int phi = 0; phi < win; phi++) {
      int k = phi - phase; if (k < 0) k += win;
      sum = __dadd_rn(sum, (double)segm_load(p, c, t + k));
    }
  }
  const double mean = __ddiv_rn(sum, (double)matFull);
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const int c = ch + chanOff;
    for (int phi = 0; phi < win; phi++) {
      int k = phi - phase; if (k < 0) k += win;
      double d = __dsub_rn((double)segm_load(p, c, t + k), mean);
      sum = __dadd_rn(sum, __dmul_rn(d, d));
    }
  }
  const double stdDev = __dsqrt_rn(__ddiv_rn(sum, (double)matFull));
  const double add = -mean;
  const int matSize = numChannels * H;
  sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const int c = ch + chanOff;
    for (int i = 0; i < H; i++) {
      double a = __dadd_rn((double)segm_load(p, c, t + i), add);
      double b = __dadd_rn((double)segm_load(p, c, t + H + i), add);
      sum = __dadd_rn(sum, __dmul_rn(a, b));
    }
  }
  return (float)__ddiv_rn(sum, __dmul_rn(__dmul_rn(stdDev, stdDev), (double)matSize));
}

__global__ void k_segm_curve(const SegmParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nOff) return;
  const float temporal = p.weight > 0.f ? segm_correlate_half(p, t, 0, 1) : 0.f;
  const float spectral = p.weight < 1.f ? segm_correlate_half(p, t, 1, p.numCh - 1) : 0.f;
  p.curve[t] = __fadd_rn(__fmul_rn(temporal, p.weight), __fmul_rn(spectral, __fsub_rn(1.0f, p.weight)));
}

// SortedSet[Break](BreakMaxOrd) + addBreak
struct BreakMachine {
  sgz_break *e;     // ascending Float.compare order on sim, unique sims
  int n, numBreaks;
  int hasLast;
  sgz_break last;
  int64_t minSpacing;

  __host__ __device__ void reset(sgz_break *store, int nb, int64_t ms) {
    e = store; n = 0; numBreaks = nb; hasLast = 0; minSpacing = ms;
    last = sgz_break{0.f, 0, 0};
  }
  __host__ __device__ bool has_space() const { return n < numBreaks; }                  // :58
  __host__ __device__ float highest() const { return n > 0 ? e[n - 1].sim : 0.f; }     // :60-62
  __host__ __device__ int find(float sim, bool &found) const {
    found = false;
    int i = 0;
    for (; i < n; i++) {
      int c = jfloat_compare(sim, e[i].sim);
      if (c == 0) { found = true; return i; }
      if (c < 0) return i;
    }
    return i;
  }
  __host__ __device__ void set_add(const sgz_break &b) {
    bool found;
    int i = find(b.sim, found);
    if (found) return;
    for (int k = n; k > i; k--) e[k] = e[k - 1];
    e[i] = b;
    n++;
  }
  __host__ __device__ void set_remove(float sim) {
    bool found;
    int i = find(sim, found);
    if (!found) return;
    for (int k = i; k + 1 < n; k++) e[k] = e[k + 1];
    n--;
  }
  __host__ __device__ void add(const sgz_break &b) {                                    // :68-83
    if (hasLast && (b.pos - last.pos) < minSpacing) {
      if (last.sim > b.sim) {
        set_remove(last.sim);
        set_add(b);
        last = b;
      }
    } else {
      set_add(b);
      if (n > numBreaks) n--;
      last = b;
      hasLast = 1;
    }
  }
};

struct PickParams {
  const float *curve;
  int nOff;
  int afStart, H, step;
  int numBreaks;
  int64_t minSpacing;
  sgz_break *out;   // [numBreaks + 1]
  int *count;
  int inShared;     // the sorted set lives in shared memory during the replay (numBreaks + 1 entries fit), else in `out`
};

// One warp walks the curve; lanes test 32 offsets per step and jump to the first state change.  The replay is a chain of
// dependent steps, so what matters is the latency of one step: the curve is staged through shared memory in chunks
// (coalesced, many loads in flight), the sorted set lives in shared memory, and ALL lanes run the (identical) state
// machine, so that no state has to be broadcast after a change.
constexpr int kPickChunk = 8192;

__global__ void k_segm_pick(const PickParams p) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  extern __shared__ __align__(16) unsigned char pickSmem[];
  float *chunk = reinterpret_cast<float *>(pickSmem);                                    // [kPickChunk]
  sgz_break *store = p.inShared ? reinterpret_cast<sgz_break *>(pickSmem + kPickChunk * sizeof(float)) : p.out;
  BreakMachine mc;
  mc.reset(store, p.numBreaks, p.minSpacing);
  int pos = 0;
  for (int base = 0; base < p.nOff; base += kPickChunk) {
    const int end = min(base + kPickChunk, p.nOff);
    __syncwarp();
    for (int i = base + lane; i < end; i += 32) chunk[i - base] = p.curve[i];
    __syncwarp();
    while (pos < end) {
      const int t = pos + lane;
      const bool active = t < end;
      const float s = active ? chunk[t - base] : 0.f;
      bool change = false;
      if (active) {
        const bool accept = mc.has_space() || s < mc.highest();
        const int64_t bpos = (int64_t)(p.afStart + t + p.H) * p.step;
        const bool collapse = mc.hasLast && (bpos - mc.last.pos) < p.minSpacing;
        change = accept && (collapse ? (mc.last.sim > s) : true);
      }
      const unsigned mask = __ballot_sync(full, change);
      if (mask == 0u) { pos += 32; continue; }
      const int l = __ffs(mask) - 1;
      const int ts = pos + l;
      const float ss = __shfl_sync(full, s, l);
      mc.add(sgz_break{ss, 0, (int64_t)(p.afStart + ts + p.H) * p.step});   // same arguments, same writes on every lane
      __syncwarp();
      pos = ts + 1;
    }
  }
  if (p.inShared) {
    __syncwarp();
    for (int i = lane; i < mc.n; i += 32) p.out[i] = store[i];
  }
  if (lane == 0) *p.count = mc.n;
}

// The same replay with the sorted set held IN REGISTERS, one entry per lane (numBreaks <= 31): find = two ballots over
// the total-order keys, insert / remove = one shuffle of every field.  A state change then costs a few dozen cycles
// instead of a chain of dependent shared-memory accesses (about every second offset of a smooth curve is a change).
__device__ __forceinline__ uint32_t segm_jkey(float x) {   // java.lang.Float.compare as an unsigned key
  return float_order_key(x != x ? __int_as_float(0x7fc00000) : x);
}

__global__ void k_segm_pick_warp(const PickParams p) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  extern __shared__ __align__(16) unsigned char pickSmem[];
  float *chunk = reinterpret_cast<float *>(pickSmem);                                    // [kPickChunk]
  uint32_t eKey = 0;      // entry `lane` of the set: ascending Float.compare order on sim, unique sims
  float eSim = 0.f;
  long long ePos = 0;
  int n = 0, hasLast = 0;
  float lastSim = 0.f, high = 0.f;
  long long lastPos = 0;
  auto set_add = [&](float sim, long long pos) {
    const uint32_t k = segm_jkey(sim);
    const unsigned lt = __ballot_sync(full, lane < n && eKey < k), eq = __ballot_sync(full, lane < n && eKey == k);
    const uint32_t uKey = __shfl_up_sync(full, eKey, 1);
    const float uSim = __shfl_up_sync(full, eSim, 1);
    const long long uPos = __shfl_up_sync(full, ePos, 1);
    if (eq != 0u) return;                                  // TreeSet.+ keeps the existing element
    const int idx = __popc(lt);
    if (lane > idx && lane <= n) { eKey = uKey; eSim = uSim; ePos = uPos; }
    if (lane == idx) { eKey = k; eSim = sim; ePos = pos; }
    n++;
  };
  auto set_remove = [&](float sim) {
    const uint32_t k = segm_jkey(sim);
    const unsigned lt = __ballot_sync(full, lane < n && eKey < k), eq = __ballot_sync(full, lane < n && eKey == k);
    const uint32_t dKey = __shfl_down_sync(full, eKey, 1);
    const float dSim = __shfl_down_sync(full, eSim, 1);
    const long long dPos = __shfl_down_sync(full, ePos, 1);
    if (eq == 0u) return;
    const int idx = __popc(lt);
    if (lane >= idx && lane < n - 1) { eKey = dKey; eSim = dSim; ePos = dPos; }
    n--;
  };
  int pos = 0;
  for (int base = 0; base < p.nOff; base += kPickChunk) {
    const int end = min(base + kPickChunk, p.nOff);
    __syncwarp();
    for (int i = base + lane; i < end; i += 32) chunk[i - base] = p.curve[i];
    __syncwarp();
    while (pos < end) {
      const int t = pos + lane;
      const bool active = t < end;
      const float s = active ? chunk[t - base] : 0.f;
      bool change = false;
      if (active) {
        const bool accept = n < p.numBreaks || s < high;                                  // :58-62, :120
        const long long bpos = (long long)(p.afStart + t + p.H) * p.step;
        const bool collapse = hasLast && (bpos - lastPos) < p.minSpacing;
        change = accept && (collapse ? (lastSim > s) : true);
      }
      const unsigned mask = __ballot_sync(full, change);
      if (mask == 0u) { pos += 32; continue; }
      const int l = __ffs(mask) - 1;
      const int ts = pos + l;
      const float bs = __shfl_sync(full, s, l);
      const long long bp = (long long)(p.afStart + ts + p.H) * p.step;
      // addBreak, :68-83
      if (hasLast && (bp - lastPos) < p.minSpacing) {
        if (lastSim > bs) {
          set_remove(lastSim);
          set_add(bs, bp);
          lastSim = bs; lastPos = bp;
        }
      } else {
        set_add(bs, bp);
        if (n > p.numBreaks) n--;
        lastSim = bs; lastPos = bp; hasLast = 1;
      }
      high = n > 0 ? __shfl_sync(full, eSim, n - 1) : 0.f;
      pos = ts + 1;
    }
  }
  if (lane < n) p.out[lane] = sgz_break{eSim, 0, (int64_t)ePos};
  if (lane == 0) *p.count = n;
}

}  // namespace sgz
