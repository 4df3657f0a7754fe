// punchout.cuh -- punch-in + punch-out search: exact replay of loop C
// (FeatureCorrelationImpl.scala:250-393) on top of the two K1 curves.
//
// Reference semantics per DB file (N frames):
//   loop A  in-curve at t in [0, nA), nA = N - minPunch - W_in + 1; tInOff = first t with
//           (entryHasSpace || simIn[t] > lowestSim)                                       (:212-223)
//   loop B  out-curve from frame tInOff + minPunch; only if N - (tInOff+minPunch) >= W_out (:260-263)
//   loop C  for pi = tInOff ..: low/hs cached; gate inSim > low*low (:342); cells
//           po = pi+minPunch .. min(pi+maxPunch, N-W_out): sim = sqrt(inSim*outSim).toFloat (:370),
//           accepted when hs || sim > low, then addMatch and low/hs refreshed        (:372-379)
//   Q5: cells the reference reads past the written part of its temp files are "not
//   searched": pi < nA and po <= N - W_out.
//
// Filling rounds (allPrio not full): one block per file; all warps pre-test and stage 2048 rows at a
// time, warp 0 replays the marked rows, 256 cells per step while no cell can change the machine state.
// Full rounds: the GPU emits (a) every t with simIn[t] > theta and (b) every cell of a gated row with
// sim > theta, theta = allPrio.last.sim at the start of the round (a provable lower bound of every
// later lowestSim); the host replays the few records in (file, pi, po) order.
#pragma once
#include "common.cuh"
#include "corr.cuh"
#include "select.cuh"

namespace sgz {

struct FillPoParams {
  const float *simIn, *simOut, *rowMax;
  BoostSrc boostIn, boostOut;
  const int64_t *fileStart;
  const int32_t *files;
  int numJobs;
  int Win, Wout, minPunchF, maxPunchF;
  int numPerFile, maxEntrySz;
  int64_t minSpacing;
  int step;
  EntryRec *entries;   // [numJobs][numPerFile + 1]
  int32_t *counts;
  // lowestSim falls back to allPrio.last.sim while entryPrio is empty (FeatureCorrelationImpl.scala:125-129), and the row
  // gate `inSim > low * low` (:342) applies even while the entry has space.  Files of one filling round run in parallel
  // with the allPrio state of the ROUND START as the guess; `meta` tells the merge whether the guess mattered:
  // meta[job] = {max in-sim of the rows (with cells) skipped while the entry was empty, in-sim of the first accepted row,
  // that row (-1: none)}.  The result is the reference's for every allPrio.last whose square lies in [meta.x, meta.y).
  int allNonEmpty;
  float allLast;
  float4 *meta;        // [numJobs]
  int staged;          // the chunk's curves fit the dynamic shared memory of the launch
  int entSmem;         // entryPrio in shared memory, behind the curves
  int curveFloats;     // floats of dynamic shared memory taken by the curves (0 when not staged)
  int prof;            // SGZ_FILL_PROF: per-file counters on stdout (developer knob)
};

__host__ __device__ __forceinline__ int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }

__device__ __forceinline__ float cell_sim(float inSim, float outSim) {
  return (float)sqrt((double)__fmul_rn(inSim, outSim));   // math.sqrt(inSim * outSim).toFloat, Float product
}

// Smallest Float product whose cell sim exceeds `a`: the Double square root and the rounding to Float are monotone,
// so `cell_sim(in, out) > a` <=> `__fmul_rn(in, out) >= cell_prod_threshold(a)` -- the replay and the candidate scan
// test a cell with one Float multiplication instead of a Double square root.  NaN = no product qualifies.
__device__ __forceinline__ float sim_of_prod(float prod) { return (float)sqrt((double)prod); }
__device__ __noinline__ float cell_prod_threshold(float a) {
  if (a != a || a == INFINITY) return NAN;      // `sim > NaN`, `sim > +inf` never hold
  if (a < 0.f) return 0.f;                      // every sim that is not NaN (products >= 0, -0.0 included)
  float pr = (float)((double)a * (double)a);    // near the boundary; +inf when a * a overflows
  for (;;) {
    if (sim_of_prod(pr) > a) {
      const float below = __int_as_float(__float_as_int(pr) - 1);   // pr > 0 here: sim_of_prod(0) = 0 is not > a >= 0
      if (!(sim_of_prod(below) > a)) return pr;
      pr = below;
    } else {
      pr = __int_as_float(__float_as_int(pr) + 1);                   // pr is finite here
    }
  }
}

// rowMax[g] = max over the grid row of punch-in offset g of simOut (NaN ignored, -inf if the row has no cell).
// cell_sim(in, out) is monotone in `out` for in > 0, so cell_sim(in, rowMax) is the exact row maximum: when it
// cannot beat `low` (and the entry has no space) no cell of the row can be accepted -> the row is skipped.
// This is the pruning the reference notes as missing (FeatureCorrelationImpl.scala:345-349), done exactly.
__global__ void k_row_max_out(const float *__restrict__ simOut, const int64_t *__restrict__ fileStart, int numFiles,
                              int64_t usedFrames, int Win, int Wout, int minPunchF, int maxPunchF,
                              float *__restrict__ rowMax) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= usedFrames) return;
  int lo = 0, hi = numFiles;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (fileStart[mid] <= g) lo = mid; else hi = mid;
  }
  const int64_t fs = fileStart[lo], N = fileStart[lo + 1] - fs, pi = g - fs;
  const int64_t nA = N - minPunchF - Win + 1;
  float m = -INFINITY;
  if (pi < nA) {
    const int64_t p0 = pi + minPunchF, p1 = i64min(pi + maxPunchF, N - Wout);
    const float *s = simOut + fs;
    for (int64_t po = p0; po <= p1; po++) m = fmaxf(m, s[po]);
  }
  rowMax[g] = m;
}

// The same maxima with a two-level scan: a block stages the out-curve under its 256 rows in shared memory together
// with the maxima of aligned 32-value groups, so that a row of 604 cells costs ~19 group maxima + two ragged ends
// instead of 604 values.  (fmaxf is order independent up to the sign of a zero maximum, which no comparison sees.)
constexpr int kRowMaxThreads = 256;
__host__ __device__ inline int row_max_tile_len(int span) { return (kRowMaxThreads + span + 32 + 31) & ~31; }
inline size_t row_max_smem_bytes(int span) { const int l = row_max_tile_len(span); return (size_t)(l + l / 32) * sizeof(float); }

__global__ void __launch_bounds__(kRowMaxThreads)
k_row_max_out_tiled(const float *__restrict__ simOut, const int64_t *__restrict__ fileStart, int numFiles,
                    int64_t usedFrames, int Win, int Wout, int minPunchF, int maxPunchF, float *__restrict__ rowMax) {
  extern __shared__ __align__(16) float rmSmem[];
  const int tileLen = row_max_tile_len(maxPunchF - minPunchF + 1);
  float *vals = rmSmem, *gmax = rmSmem + tileLen;
  const int64_t g0 = blockIdx.x * (int64_t)kRowMaxThreads;
  const int64_t tLo = (g0 + minPunchF) & ~(int64_t)31;                                   // aligned tile start
  const int64_t tHi = i64min(g0 + kRowMaxThreads - 1 + maxPunchF, usedFrames - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < tileLen; i += kRowMaxThreads) vals[i] = tLo + i <= tHi ? simOut[tLo + i] : -INFINITY;
  __syncthreads();
  for (int b = warp; b < tileLen / 32; b += kRowMaxThreads / 32) {
    float m = fmaxf(-INFINITY, vals[32 * b + lane]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) gmax[b] = m;
  }
  __syncthreads();
  const int64_t g = g0 + threadIdx.x;
  if (g >= usedFrames) return;
  int lo = 0, hi = numFiles;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (fileStart[mid] <= g) lo = mid; else hi = mid;
  }
  const int64_t fs = fileStart[lo], N = fileStart[lo + 1] - fs, pi = g - fs;
  const int64_t nA = N - minPunchF - Win + 1;
  float m = -INFINITY;
  if (pi < nA) {
    int i = (int)(fs + pi + minPunchF - tLo);
    const int last = (int)(fs + i64min(pi + maxPunchF, N - Wout) - tLo);
    for (; i <= last && (i & 31); i++) m = fmaxf(m, vals[i]);
    for (; i + 31 <= last; i += 32) m = fmaxf(m, gmax[i >> 5]);
    for (; i <= last; i++) m = fmaxf(m, vals[i]);
  }
  rowMax[g] = m;
}

inline cudaError_t launch_row_max_out(cudaStream_t st, const float *simOut, const int64_t *fileStart, int numFiles,
                                      int64_t usedFrames, int Win, int Wout, int minPunchF, int maxPunchF, float *rowMax) {
  const int span = maxPunchF - minPunchF + 1;
  const unsigned blocks = (unsigned)ceil_div<int64_t>(usedFrames, kRowMaxThreads);
  const size_t smem = span > 0 ? row_max_smem_bytes(span) : 0;
  if (span > 64 && smem <= 200 * 1024 && getenv("SGZ_PO_GLOBAL") == nullptr) {   // developer knob: the untiled kernels
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_row_max_out_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return e;
    }
    k_row_max_out_tiled<<<blocks, kRowMaxThreads, smem, st>>>(simOut, fileStart, numFiles, usedFrames, Win, Wout, minPunchF,
                                                              maxPunchF, rowMax);
  } else {
    k_row_max_out<<<blocks, kRowMaxThreads, 0, st>>>(simOut, fileStart, numFiles, usedFrames, Win, Wout, minPunchF,
                                                     maxPunchF, rowMax);
  }
  return cudaGetLastError();
}

// One block per file.  The replay itself is sequential (warp 0), but while the entry has no space `lowest` never
// decreases (addMatch only replaces or drops lower sims), so "could this row pass the gate" evaluated with the
// `low` of the chunk start is a superset of the rows that pass later: all warps test a chunk of kFillChunk rows
// in parallel and leave one bit per row, warp 0 walks only the marked rows and re-evaluates them exactly.  Should
// the entry regain space inside a chunk (a sim equal to a stored one collapsing into the last match) the marks
// are dropped until the next chunk.  While they test, the warps also stage what the chunk's rows can touch -- the
// in-curve and row maxima of the chunk and the kFillChunk + span out-curve values under its grid rows -- in shared
// memory (`staged`; grids too wide for it are read from global memory), so that the replay waits for memory only
// when a cell is accepted (the two boosts).  Cells are tested through cell_prod_threshold, the Double square root
// is taken for accepted cells only.
constexpr int kFillChunk = 2048;   // rows per parallel pre-test
constexpr int kFillThreads = 256;
constexpr int kFillPerThread = kFillChunk / kFillThreads;
constexpr size_t kFillSmemMax = 200 * 1024;
__host__ __device__ inline size_t fill_po_smem_bytes(int span) { return (size_t)(3 * kFillChunk + span) * sizeof(float); }
constexpr int kFillCellSteps = 8;  // 32-cell steps per test of the replay

__global__ void __launch_bounds__(kFillThreads) k_replay_fill_po(const FillPoParams p) {
  __shared__ unsigned marks[kFillChunk / 32];
  __shared__ float sLow;
  __shared__ int sHs;
  extern __shared__ __align__(16) float fillSmem[];
  float *sIn = fillSmem, *sRm = fillSmem + kFillChunk, *sOut = fillSmem + 2 * kFillChunk;   // sOut: [kFillChunk + span]
  const bool staged = p.staged != 0;
  const int job = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const int f = p.files[job];
  const int64_t fs = p.fileStart[f];
  const int64_t N = p.fileStart[f + 1] - fs;
  const int64_t nA = N - p.minPunchF - p.Win + 1;
  const int64_t poMax = N - p.Wout;                       // last frame with an out-correlation
  const int span = p.maxPunchF - p.minPunchF + 1;
  // entryPrio lives in shared memory during the replay (warp 0, lane 0) unless numPerFile is huge
  EntryRec *gEnt = p.entries + (size_t)job * (p.numPerFile + 1);
  EntryRec *ent = p.entSmem ? reinterpret_cast<EntryRec *>(fillSmem + p.curveFloats) : gEnt;
  Machine mc;
  mc.reset(ent, p.numPerFile, p.maxEntrySz, p.allNonEmpty, p.allLast, p.minSpacing, p.step);
  bool entEmpty = true;             // entryPrio still empty: `low` is allPrio.last.sim (or 0)
  float metaM = -INFINITY, metaV = 0.f;
  int metaRow = -1;
  // filling round: entryHasSpace holds at the file start, so tInOff = 0; loop B needs one full window
  const bool any = nA > 0 && (N - p.minPunchF) >= p.Wout && span > 0 && p.maxEntrySz > 0;
  bool hs = mc.has_space();
  float low = mc.lowest();
  int hasLast = 0;
  float lastSim = 0.f;
  int lastStop = 0;
  float tLow = 0.f, tLast = 0.f;    // cell_prod_threshold of `low` and of the last match's sim (warp 0)
  float tLowOf = 0.f;               // the `low` tLow was computed for
  int64_t pi = 0;
  long long tMark = 0, tWalk = 0, tCells = 0, t0 = 0, t1 = 0;
  int nGated = 0, nChanges = 0, nSteps = 0;
  if (threadIdx.x == 0) { sLow = low; sHs = hs ? 1 : 0; }
  for (int64_t chunk = 0; any && chunk < nA; chunk += kFillChunk) {
    if (p.prof) t0 = clock64();
    __syncthreads();
    const float lowC = sLow;
    const bool hsC = sHs != 0;
    const float tLowC = cell_prod_threshold(lowC);
    // ---- all warps: which rows of the chunk could pass the gate (state of the chunk start) ----
    const int64_t chunkEnd = chunk + kFillChunk < nA ? chunk + kFillChunk : nA;
    float inR[kFillPerThread], rmR[kFillPerThread];
#pragma unroll
    for (int j = 0; j < kFillPerThread; j++) {      // warp w looks at the 32-row groups w, w + nWarps, ...
      const int64_t r = chunk + (int64_t)j * kFillThreads + threadIdx.x;
      inR[j] = r < nA ? p.simIn[fs + r] : 0.f;
      rmR[j] = r < nA ? p.rowMax[fs + r] : 0.f;
    }
    if (staged) {
      const int64_t outLo = chunk + p.minPunchF, outHi = i64min(chunkEnd - 1 + p.maxPunchF, poMax);
      for (int64_t i = threadIdx.x; outLo + i <= outHi; i += kFillThreads) sOut[i] = p.simOut[fs + outLo + i];
    }
#pragma unroll
    for (int j = 0; j < kFillPerThread; j++) {
      const int i = j * kFillThreads + threadIdx.x;
      const int64_t r = chunk + i;
      if (staged) { sIn[i] = inR[j]; sRm[i] = rmR[j]; }
      bool could = false;
      if (r < nA) {
        const int64_t cellsR = i64min(poMax - (r + p.minPunchF) + 1, (int64_t)span);
        if (hsC) could = cellsR > 0;
        else could = cellsR > 0 && inR[j] > __fmul_rn(lowC, lowC) && __fmul_rn(inR[j], rmR[j]) >= tLowC;
      }
      const unsigned m = __ballot_sync(full, could);
      if (lane == 0) marks[i >> 5] = m;
    }
    __syncthreads();
    if (warp != 0) continue;
    if (p.prof) { t1 = clock64(); tMark += t1 - t0; }
    // ---- warp 0: exact replay over the marked rows ----
    bool marksValid = !hs;
    static_assert(kFillChunk / 32 == 64, "markedGroups is one 64-bit word");
    const unsigned long long markedGroups = ((unsigned long long)__ballot_sync(full, marks[32 + lane] != 0u) << 32) |
                                            __ballot_sync(full, marks[lane] != 0u);
    if (chunk == 0) { tLow = cell_prod_threshold(low); tLowOf = low; tLast = cell_prod_threshold(lastSim); }
    if (pi < chunk) pi = chunk;
    while (pi < chunkEnd) {
      if (marksValid) {                              // jump to the next marked row
        const int o = (int)(pi - chunk);
        int g = o >> 5;
        unsigned m = marks[g] & (full << (o & 31));
        if (m == 0u) {
          const unsigned long long later = g < 63 ? markedGroups >> (g + 1) << (g + 1) : 0ull;
          if (later == 0ull) { pi = chunkEnd; continue; }
          g = __ffsll((long long)later) - 1;
          m = marks[g];
        }
        pi = chunk + 32 * g + (__ffs(m) - 1);
      }
      // rows pi .. pi+31: the first one whose gate holds (state is constant while nothing is accepted)
      const int64_t r = pi + lane;
      const bool inChunk = r < chunkEnd;             // rows of the next chunk wait for their own pre-test
      const float in = inChunk ? (staged ? sIn[r - chunk] : p.simIn[fs + r]) : 0.f;
      const float rm = inChunk ? (staged ? sRm[r - chunk] : p.rowMax[fs + r]) : 0.f;
      const int64_t cellsR = i64min(poMax - (r + p.minPunchF) + 1, (int64_t)span);
      bool gate = inChunk && cellsR > 0 && in > __fmul_rn(low, low);
      // exact pruning: without space in the entry a row only matters if its best cell beats `low`, and a row that
      // collapses into the last match only if its best cell beats that match
      if (gate && !hs) gate = __fmul_rn(in, rm) >= tLow;
      if (gate && hasLast && ((r - (int64_t)lastStop) * p.step < p.minSpacing)) gate = __fmul_rn(in, rm) >= tLast;
      const unsigned rmask = __ballot_sync(full, gate);
      nSteps++;
      if (entEmpty) {   // rows the gate turned away before the first match: would a lower allPrio.last have let them in?
        const bool skipped = inChunk && cellsR > 0 && !gate && (rmask == 0u || lane < __ffs(rmask) - 1);
        float m = skipped ? in : -INFINITY;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(full, m, d));
        metaM = fmaxf(metaM, m);
      }
      if (rmask == 0u) { pi = pi + 32 < chunkEnd ? pi + 32 : chunkEnd; continue; }
      const int rl = __ffs(rmask) - 1;
      const int64_t row = pi + rl;
      const float inS = __shfl_sync(full, in, rl);
      if (entEmpty) { metaRow = (int)row; metaV = inS; entEmpty = false; }   // hs holds: the row's first cell is accepted
      const int n = (int)__shfl_sync(full, (int)cellsR, rl);
      nGated++;
      const long long tc0 = p.prof ? clock64() : 0;
      const float *outRow = staged ? sOut + (row - chunk) : p.simOut + fs + row + p.minPunchF;
      // cells c, c+1, ... of the row, kFillCellSteps x 32 per test; ONE copy of the state-change code (unrolled per step
      // the kernel outgrew the instruction cache and every step paid for it)
      int c = 0;
      while (c < n) {
        float prod[kFillCellSteps];
#pragma unroll
        for (int q = 0; q < kFillCellSteps; q++) {
          const int k = c + 32 * q + lane;
          prod[q] = k < n ? __fmul_rn(inS, outRow[k]) : NAN;
        }
        const bool collapse = hasLast && ((row - (int64_t)lastStop) * p.step < p.minSpacing);
        unsigned changes = 0u;                       // bit q: cell c + 32 q + lane would change the state
#pragma unroll
        for (int q = 0; q < kFillCellSteps; q++) {
          const bool accept = hs || prod[q] >= tLow;                                          // sim > low
          const bool change = c + 32 * q + lane < n && accept && (collapse ? prod[q] >= tLast : true);   // lastSim < sim
          changes |= change ? 1u << q : 0u;
        }
        if (!__any_sync(full, changes != 0u)) { c += 32 * kFillCellSteps; continue; }
        unsigned cmask = 0u;
        int qHit = -1;
        float prodHit = 0.f;
#pragma unroll
        for (int q = 0; q < kFillCellSteps; q++) {
          const unsigned m = __ballot_sync(full, (changes >> q) & 1u);
          if (qHit < 0 && m != 0u) { qHit = q; cmask = m; prodHit = prod[q]; }
        }
        const int cl = __ffs(cmask) - 1;
        const float ss = sim_of_prod(__shfl_sync(full, prodHit, cl));
        const int cell = c + 32 * qHit + cl;
        const int64_t pos = row + p.minPunchF + cell;
        if (lane == 0) {
          EntryRec m{ss, (int32_t)row, (int32_t)pos, 0.f, 0.f};   // addMatch does not look at the boosts: filled in at the end
          mc.add(m);
          hs = mc.has_space();
          low = mc.lowest();
          hasLast = mc.hasLast;
          lastSim = mc.last.sim;
          lastStop = mc.last.stopOff;
        }
        hs = __shfl_sync(full, (int)hs, 0) != 0;
        low = __shfl_sync(full, low, 0);
        hasLast = __shfl_sync(full, hasLast, 0);
        lastSim = __shfl_sync(full, lastSim, 0);
        lastStop = __shfl_sync(full, lastStop, 0);
        if (__float_as_int(low) != __float_as_int(tLowOf)) { tLow = cell_prod_threshold(low); tLowOf = low; }
        tLast = cell_prod_threshold(lastSim);
        if (hs) marksValid = false;     // `low` may fall again: the marks are no superset any more
        c = cell + 1;
        nChanges++;
      }
      if (p.prof) tCells += clock64() - tc0;
      pi = row + 1;
    }
    if (lane == 0) { sLow = low; sHs = hs ? 1 : 0; }
    if (p.prof) tWalk += clock64() - t1;
  }
  if (warp == 0 && lane == 0) {
    p.counts[job] = mc.n;
    p.meta[job] = make_float4(metaM, metaV, __int_as_float(metaRow), 0.f);
    for (int i = 0; i < mc.n; i++) {
      EntryRec e = ent[i];
      e.boostIn = p.boostIn.at(fs + e.piOff, e.piOff);
      e.boostOut = p.boostOut.at(fs + e.stopOff, e.stopOff);
      gEnt[i] = e;
    }
  }
  if (p.prof && warp == 0 && lane == 0)
    printf("k_replay_fill_po file %d: rows %lld, row steps %d, gated rows %d, state changes %d; cycles: marks %lld, walk %lld (cells %lld)\n",
           f, (long long)nA, nSteps, nGated, nChanges, tMark, tWalk, tCells);
}

struct CandPoParams {
  const float *simIn, *simOut, *rowMax;
  BoostSrc boostIn, boostOut;
  const int64_t *fileStart;
  const int32_t *files;
  const float *thresholds;
  int numJobs;
  int Win, Wout, minPunchF, maxPunchF;
  int fileBase;
  sgz_record *out;
  int cap;
  int *counter;
};

__device__ __forceinline__ void emit_record(const CandPoParams &p, unsigned mask, int lane, bool hit, const sgz_record &r) {
  int slot0 = 0;
  if (lane == 0) slot0 = atomicAdd(p.counter, __popc(mask));
  slot0 = __shfl_sync(0xffffffffu, slot0, 0);
  if (hit) {
    const int slot = slot0 + __popc(mask & ((1u << lane) - 1u));
    if (slot < p.cap) p.out[slot] = r;
  }
}

// one block per (file, row segment); warps stride over rows
__global__ void k_candidates_po(const CandPoParams p) {
  const int job = blockIdx.x;
  const int f = p.files[job];
  const float thr = p.thresholds[job];
  const float thr2 = __fmul_rn(thr, thr);
  const float tThr = cell_prod_threshold(thr);
  const int64_t fs = p.fileStart[f];
  const int64_t N = p.fileStart[f + 1] - fs;
  const int64_t nA = N - p.minPunchF - p.Win + 1;
  const int64_t poMax = N - p.Wout;
  const int span = p.maxPunchF - p.minPunchF + 1;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  if (nA <= 0 || span <= 0) return;
  // blockIdx.y = row segment of the file (the host sorts the records, so the emission order is free)
  const int64_t segLen = ((nA + gridDim.y - 1) / gridDim.y + 31) & ~(int64_t)31;
  const int64_t r0 = (int64_t)blockIdx.y * segLen, r1 = i64min(nA, r0 + segLen);
  // (a) in-curve offsets above the threshold: decide tInOff in the replay
  for (int64_t base = r0 + (int64_t)warp * 32; base < r1; base += (int64_t)nWarps * 32) {
    const int64_t t = base + lane;
    const float s = t < r1 ? p.simIn[fs + t] : 0.f;
    const bool hit = t < r1 && s > thr;
    const unsigned mask = __ballot_sync(full, hit);
    if (mask == 0u) continue;
    sgz_record r{p.fileBase + f, 3, (int32_t)t, -1, s, 0.f, 0.f, 0};
    emit_record(p, mask, lane, hit, r);
  }
  // (b) cells of gated rows
  for (int64_t rbase = r0 + (int64_t)warp * 32; rbase < r1; rbase += (int64_t)nWarps * 32) {
    const int64_t r = rbase + lane;
    const float in = r < r1 ? p.simIn[fs + r] : 0.f;
    unsigned rmask = __ballot_sync(full, r < r1 && in > thr2 && __fmul_rn(in, p.rowMax[fs + r]) >= tThr);
    while (rmask) {
      const int rl = __ffs(rmask) - 1;
      rmask &= rmask - 1;
      const int64_t row = rbase + rl;
      const float inS = __shfl_sync(full, in, rl);
      const int n = (int)i64min(poMax - (row + p.minPunchF) + 1, (int64_t)span);
      for (int c = 0; c < n; c += 32) {
        const int k = c + lane;
        const int64_t po = row + p.minPunchF + k;
        const float prod = k < n ? __fmul_rn(inS, p.simOut[fs + po]) : NAN;
        const bool hit = prod >= tThr;                                                    // sim > thr
        const unsigned mask = __ballot_sync(full, hit);
        if (mask == 0u) continue;
        const float s = hit ? sim_of_prod(prod) : 0.f;
        sgz_record rec{p.fileBase + f, 0, (int32_t)row, (int32_t)po, s, hit ? p.boostIn.at(fs + row, row) : 0.f,
                       hit ? p.boostOut.at(fs + po, po) : 0.f, __float_as_int(inS)};
        emit_record(p, mask, lane, hit, rec);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host: one selection round / merge in punch-out mode
// ---------------------------------------------------------------------------------------------
inline int corr_select_punchout(sgz_corr *job, int32_t *nRecords) {
  sgz_ctx *ctx = job->ctx;
  sgz_db *db = job->db;
  const int K = job->cfg.numMatches, npf = job->cfg.numPerFile;
  const int room = K - (int)job->allPrio.size();
  const int myLo = job->myFirst, myHi = job->myFirst + db->numFiles();
  if (room > 0) {
    const int nb = room >= npf ? room / npf : 1;
    const int m = room >= npf ? npf : room;
    job->roundKind = 0;
    job->roundFirst = job->nextFile;
    job->roundCount = std::min(nb, job->nFilesGlobal - job->nextFile);
    job->roundMaxEntrySz = m;
    std::vector<int32_t> files;
    for (int g = job->roundFirst; g < job->roundFirst + job->roundCount; g++)
      if (g >= myLo && g < myHi) files.push_back(g - myLo);
    if (!files.empty()) {
      const int nj = (int)files.size();
      SGZ_TRY(job->dFiles.alloc(nj));
      SGZ_TRY(job->dCounts.alloc(nj));
      SGZ_TRY(job->dEntries.alloc((size_t)nj * (npf + 1)));
      SGZ_CUDA(cudaMemcpyAsync(job->dFiles.p, files.data(), nj * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      FillPoParams fp{};
      fp.simIn = job->simIn.p; fp.simOut = job->simOut.p;
      fp.boostIn = boost_src(job, job->qin, job->boostIn.p); fp.boostOut = boost_src(job, job->qout, job->boostOut.p);
      fp.rowMax = job->rowMaxOut.p;
      fp.fileStart = db->dFileStart.p; fp.files = job->dFiles.p; fp.numJobs = nj;
      fp.Win = job->qin.W; fp.Wout = job->qout.W; fp.minPunchF = job->minPunchF; fp.maxPunchF = job->maxPunchF;
      fp.numPerFile = npf; fp.maxEntrySz = m; fp.minSpacing = job->cfg.minSpacing; fp.step = job->step;
      fp.entries = job->dEntries.p; fp.counts = job->dCounts.p;
      SGZ_TRY(job->dMeta.alloc(nj));
      fp.meta = job->dMeta.p;
      fp.allNonEmpty = job->allPrio.empty() ? 0 : 1;
      fp.allLast = job->allPrio.empty() ? 0.f : job->allPrio.back().sim;
      static const bool fillProf = getenv("SGZ_FILL_PROF") != nullptr;
      fp.prof = fillProf ? 1 : 0;
      SGZ_TRY(ctx->begin_call());
      const size_t curveBytes = fill_po_smem_bytes(fp.maxPunchF - fp.minPunchF + 1);
      const size_t entBytes = (size_t)(npf + 1) * sizeof(EntryRec);
      const bool poGlobal = getenv("SGZ_PO_GLOBAL") != nullptr;   // developer knob: curves and entries stay in global memory
      fp.staged = fp.maxPunchF >= fp.minPunchF && curveBytes <= kFillSmemMax && !poGlobal ? 1 : 0;
      fp.curveFloats = fp.staged ? (int)(curveBytes / sizeof(float)) : 0;
      fp.entSmem = entBytes <= 16 * 1024 && !poGlobal ? 1 : 0;
      const size_t fillSmem = (fp.staged ? curveBytes : 0) + (fp.entSmem ? entBytes : 0);
      if (fillSmem > 48 * 1024)
        SGZ_CUDA(cudaFuncSetAttribute(k_replay_fill_po, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(kFillSmemMax + 16 * 1024)));
      k_replay_fill_po<<<(unsigned)nj, kFillThreads, fillSmem, ctx->stream>>>(fp);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_TRY(ctx->end_call());
      job->selectMs += ctx->lastMs;
      std::vector<int32_t> counts(nj);
      std::vector<EntryRec> ents((size_t)nj * (npf + 1));
      std::vector<float4> meta((size_t)nj);
      SGZ_CUDA(cudaMemcpyAsync(counts.data(), job->dCounts.p, nj * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(ents.data(), job->dEntries.p, ents.size() * sizeof(EntryRec), cudaMemcpyDeviceToHost,
                               ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(meta.data(), job->dMeta.p, nj * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int j = 0; j < nj; j++) {
        // kind 4 = "gate interval" of the file (one per file of the round, before its matches): piOff = first accepted
        // row or -1, sim = its in-sim, boostIn = largest in-sim the gate turned away before it
        int32_t row;
        memcpy(&row, &meta[j].z, 4);
        job->localRecords.push_back(sgz_record{myLo + files[j], 4, row, -1, meta[j].y, meta[j].x, 0.f, 0});
        for (int k = 0; k < counts[j]; k++) {
          const EntryRec &e = ents[(size_t)j * (npf + 1) + k];
          job->localRecords.push_back(sgz_record{myLo + files[j], 1, e.piOff, e.stopOff, e.sim, e.boostIn, e.boostOut, 0});
        }
      }
    }
  } else {
    // full round over a geometrically growing batch of files; theta = allPrio.last.sim now
    job->roundKind = 1;
    job->roundFirst = job->nextFile;
    const int batch = std::max(64, job->nextFile);
    job->roundCount = std::min(batch, job->nFilesGlobal - job->nextFile);
    const float theta = job->allPrio.back().sim;
    std::vector<int32_t> files;
    std::vector<float> thr;
    if (theta == theta) {   // allPrio.last NaN: `sim > NaN` never holds
      for (int g = job->roundFirst; g < job->roundFirst + job->roundCount; g++) {
        if (g < myLo || g >= myHi) continue;
        const sgz_file_summary &s = job->globalSummary[g];
        // no cell can exceed sqrt(maxIn * maxOut); in-curve must exceed theta somewhere to open tIn
        // (cell sims are >= 0 or NaN; a non-positive in-curve never passes the gate inSim > low*low)
        const float ub = s.maxSim > 0.f
                             ? (s.maxSimOut > 0.f ? (float)sqrt((double)s.maxSim * (double)s.maxSimOut) * 1.000001f : 0.f)
                             : -INFINITY;
        if (s.maxSim > theta && ub > theta) { files.push_back(g - myLo); thr.push_back(theta); }
      }
    }
    if (!files.empty()) {
      const int nj = (int)files.size();
      SGZ_TRY(job->dFiles.alloc(nj));
      SGZ_TRY(job->dThr.alloc(nj));
      SGZ_TRY(job->dCounter.alloc(1));
      SGZ_CUDA(cudaMemcpyAsync(job->dFiles.p, files.data(), nj * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(job->dThr.p, thr.data(), nj * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      int cap = std::max(1 << 16, (int)job->dRecs.n);
      for (;;) {
        SGZ_TRY(job->dRecs.alloc(cap));
        SGZ_CUDA(cudaMemsetAsync(job->dCounter.p, 0, sizeof(int), ctx->stream));
        CandPoParams cp{};
        cp.simIn = job->simIn.p; cp.simOut = job->simOut.p;
        cp.boostIn = boost_src(job, job->qin, job->boostIn.p); cp.boostOut = boost_src(job, job->qout, job->boostOut.p);
        cp.rowMax = job->rowMaxOut.p;
        cp.fileStart = db->dFileStart.p; cp.files = job->dFiles.p; cp.thresholds = job->dThr.p; cp.numJobs = nj;
        cp.Win = job->qin.W; cp.Wout = job->qout.W; cp.minPunchF = job->minPunchF; cp.maxPunchF = job->maxPunchF;
        cp.fileBase = myLo; cp.out = job->dRecs.p; cp.cap = cap; cp.counter = job->dCounter.p;
        SGZ_TRY(ctx->begin_call());
        k_candidates_po<<<dim3((unsigned)nj, (unsigned)std::max(1, std::min(16, 4 * ctx->smCount / nj))), 256, 0, ctx->stream>>>(cp);
        SGZ_LAUNCH_CHECK(ctx);
        SGZ_TRY(ctx->end_call());
        job->selectMs += ctx->lastMs;
        int count = 0;
        SGZ_CUDA(cudaMemcpyAsync(&count, job->dCounter.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (count <= cap) {
          job->localRecords.resize(count);
          if (count > 0) {
            SGZ_CUDA(cudaMemcpyAsync(job->localRecords.data(), job->dRecs.p, (size_t)count * sizeof(sgz_record),
                                     cudaMemcpyDeviceToHost, ctx->stream));
            SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
          }
          break;
        }
        cap = count + count / 2;
      }
    }
  }
  *nRecords = (int32_t)job->localRecords.size();
  return SGZ_OK;
}

inline int corr_merge_punchout(sgz_corr *job, const sgz_record *all, int32_t nAll, int32_t *done) {
  const int K = job->cfg.numMatches, npf = job->cfg.numPerFile, step = job->step;
  std::vector<sgz_record> recs(all, all + nAll);
  auto merge_entry = [&](const EntryRec *e, int n, int file) {
    for (int i = 0; i < n; i++) {
      sgz_match m;
      m.sim = e[i].sim; m.file = file;
      m.start = feat_to_full(e[i].piOff, step);
      m.stop = feat_to_full(e[i].stopOff, step);
      m.boostIn = e[i].boostIn; m.boostOut = e[i].boostOut;
      allprio_add(job->allPrio, m);
    }
    if ((int)job->allPrio.size() > K) job->allPrio.resize(K);
  };
  if (job->roundKind == 0) {
    std::stable_sort(recs.begin(), recs.end(), [](const sgz_record &a, const sgz_record &b) { return a.file < b.file; });
    size_t i = 0;
    std::vector<EntryRec> ent;
    for (int g = job->roundFirst; g < job->roundFirst + job->roundCount; g++) {
      ent.clear();
      while (i < recs.size() && recs[i].file < g) i++;
      bool valid = true;
      for (; i < recs.size() && recs[i].file == g; i++) {
        if (recs[i].kind == 1) ent.push_back(EntryRec{recs[i].sim, recs[i].piOff, recs[i].poOff, recs[i].boostIn, recs[i].boostOut});
        else if (recs[i].kind == 4) {
          // the file ran with the allPrio of the round start; with the allPrio of NOW (files before g merged) the row
          // gate `inSim > low * low` must pick the same first row, else the file is replayed as the head of the next round
          const float low = job->allPrio.empty() ? 0.f : job->allPrio.back().sim;
          const float thr = low * low;
          valid = !(recs[i].boostIn > thr) && (recs[i].piOff < 0 || recs[i].sim > thr);
        }
      }
      if (!valid) {   // never the first file of a round: its guess is the true state
        job->roundCount = g - job->roundFirst;
        break;
      }
      merge_entry(ent.data(), (int)ent.size(), g);
    }
  } else {
    // (file, pi, kind 3 before cells, po)
    std::sort(recs.begin(), recs.end(), [](const sgz_record &a, const sgz_record &b) {
      if (a.file != b.file) return a.file < b.file;
      if (a.piOff != b.piOff) return a.piOff < b.piOff;
      if (a.kind != b.kind) return a.kind > b.kind;
      return a.poOff < b.poOff;
    });
    std::vector<EntryRec> store((size_t)npf + 1);
    size_t i = 0;
    while (i < recs.size()) {
      const int g = recs[i].file;
      size_t j = i;
      while (j < recs.size() && recs[j].file == g) j++;
      Machine mc;
      const int allSize = (int)job->allPrio.size();
      mc.reset(store.data(), npf, std::min(K - allSize, npf), allSize > 0, allSize > 0 ? job->allPrio.back().sim : 0.f,
               job->cfg.minSpacing, step);
      // loop A: tInOff = first t with entryHasSpace || simIn[t] > lowestSim (state of the file start)
      int tInOff = -1;
      for (size_t k = i; k < j; k++)
        if (recs[k].kind == 3 && (mc.has_space() || recs[k].sim > mc.lowest())) { tInOff = recs[k].piOff; break; }
      const sgz_file_summary &s = job->globalSummary[g];
      const int64_t N = (int64_t)s.numOffsets + job->minPunchF + job->qin.W - 1;   // nA = N - minPunch - W_in + 1
      if (tInOff >= 0 && N - (tInOff + job->minPunchF) >= job->qout.W) {
        size_t k = i;
        while (k < j) {
          if (recs[k].kind != 0 || recs[k].piOff < tInOff) { k++; continue; }
          const int row = recs[k].piOff;
          float inSim;
          memcpy(&inSim, &recs[k].aux, 4);
          float low = mc.lowest();
          bool hs = mc.has_space();
          const bool gate = inSim > low * low;                                     // :342
          for (; k < j && recs[k].piOff == row; k++) {
            if (recs[k].kind != 0 || !gate) continue;
            const sgz_record &r = recs[k];
            if (hs || r.sim > low) {
              mc.add(EntryRec{r.sim, r.piOff, r.poOff, r.boostIn, r.boostOut});
              low = mc.lowest();
              hs = mc.has_space();
            }
          }
        }
      }
      merge_entry(mc.e, mc.n, g);
      i = j;
    }
  }
  job->nextFile = job->roundFirst + job->roundCount;
  if (job->nextFile >= job->nFilesGlobal) job->finished = true;
  *done = job->finished ? 1 : 0;
  if (job->finished) job->progress = 1.0f;
  return SGZ_OK;
}

}  // namespace sgz
