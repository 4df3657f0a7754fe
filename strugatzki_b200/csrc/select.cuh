// select.cuh -- K2: top-K selection that reproduces the reference's order-dependent
// two-level priority queue EXACTLY (FeatureCorrelationImpl.scala:113-150, 212-240, 322-389,
// 399-400; ordering Api/FeatureCorrelation.scala:75-77; SpanUtil.spacing SpanUtil.scala:38-43).
//
// The reference walks every offset of every file sequentially through `addMatch`.  Observations
// that make this parallel without changing a single decision ("selection"):
//   * between two state CHANGES of the machine, whether an offset changes the state is a pure
//     predicate of (sim, offset) and the current state -> a warp evaluates 32 offsets per step
//     and jumps to the first lane whose predicate holds (ballot + ffs);
//   * once allPrio is full, entryHasSpace is false for good and an offset can only be accepted
//     if sim > allPrio.last.sim, which never decreases -> any lower bound of it is a safe
//     pre-filter; the GPU emits only those candidates and the tiny ordered replay runs on them.
#pragma once
#include "common.cuh"
#include "corr_fix.cuh"

namespace sgz {

struct EntryRec {
  float sim;
  int32_t piOff;     // punch-in offset, feature frames, file local
  int32_t stopOff;   // punch-in only: piOff + W_in;  punch-out mode: poOff
  float boostIn;
  float boostOut;
};

// The per-file state machine.  Storage for the entry set is supplied by the caller
// (capacity >= numPerFile + 1).
struct Machine {
  EntryRec *e;        // entryPrio: descending Float.compare order on sim, unique sims
  int n;
  int numPerFile;
  int maxEntrySz;     // min(numMatches - allPrio.size, numPerFile), constant within one file
  int hasLast;
  EntryRec last;      // lastEntryMatch
  int allNonEmpty;
  float allLast;      // allPrio.last.sim
  int64_t minSpacing;
  int step;

  __host__ __device__ void reset(EntryRec *store, int npf, int maxSz, int allNE, float allL, int64_t minSp,
                                 int stp) {
    e = store; n = 0; numPerFile = npf; maxEntrySz = maxSz; hasLast = 0;
    allNonEmpty = allNE; allLast = allL; minSpacing = minSp; step = stp;
    last = EntryRec{0.f, 0, 0, 0.f, 0.f};
  }
  __host__ __device__ bool has_space() const { return n < maxEntrySz; }                 // :120-123
  __host__ __device__ float lowest() const {                                            // :125-129
    if (n > 0) return e[n - 1].sim;
    if (allNonEmpty) return allLast;
    return 0.f;
  }
  // position of sim in the descending set; found = an element compares equal
  __host__ __device__ int find(float sim, bool &found) const {
    found = false;
    int i = 0;
    for (; i < n; i++) {
      int c = jfloat_compare(e[i].sim, sim);  // MatchMinOrd.compare(new, e[i]) = e[i].sim compare new.sim
      if (c == 0) { found = true; return i; }
      if (c < 0) return i;
    }
    return i;
  }
  __host__ __device__ void set_add(const EntryRec &m) {  // SortedSet + : never overwrites an equal key
    bool found;
    int i = find(m.sim, found);
    if (found) return;
    for (int k = n; k > i; k--) e[k] = e[k - 1];
    e[i] = m;
    n++;
  }
  __host__ __device__ void set_remove(float sim) {       // SortedSet - : removes the equal key
    bool found;
    int i = find(sim, found);
    if (!found) return;
    for (int k = i; k + 1 < n; k++) e[k] = e[k + 1];
    n--;
  }
  __host__ __device__ int64_t spacing_to_last(const EntryRec &m) const {  // SpanUtil.spacing(m.punch, last.punch)
    int64_t aStart = (int64_t)m.piOff * step, aStop = (int64_t)m.stopOff * step;
    int64_t bStart = (int64_t)last.piOff * step, bStop = (int64_t)last.stopOff * step;
    return aStart < bStart ? bStart - aStop : aStart - bStop;
  }
  __host__ __device__ bool collapses(const EntryRec &m) const {
    return hasLast && spacing_to_last(m) < minSpacing;
  }
  // addMatch, :135-150
  __host__ __device__ void add(const EntryRec &m) {
    if (collapses(m)) {
      if (last.sim < m.sim) {
        set_remove(last.sim);
        set_add(m);
        last = m;
      }
    } else {
      set_add(m);
      if (n > numPerFile) n--;   // entryPrio -= entryPrio.last
      last = m;
      hasLast = 1;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// filling phase, punch-in only: one block replays one file over its whole curve
// ---------------------------------------------------------------------------------------------
struct FillParams {
  const float *sim;
  BoostSrc boost;
  const int64_t *fileStart;
  const int32_t *files;     // local file indices to replay [numJobs]
  int numJobs;
  int W;
  int tailExtra;
  int numPerFile;
  int maxEntrySz;
  int64_t minSpacing;
  int step;
  EntryRec *entries;        // [numJobs][numPerFile + 1]
  int32_t *counts;          // [numJobs]
};

// One block per file.  The replay itself is sequential (warp 0), but while the entry has no space `lowest` never
// decreases (addMatch only replaces or drops lower sims), so "sim > low" evaluated with the `low` of the chunk start marks
// a superset of the offsets that can change the state later in the chunk: all warps stage a chunk of the curve in shared
// memory and leave one bit per offset, warp 0 visits only the marked 32-offset groups and re-evaluates them exactly.
// While the entry still has space (the file start, or a collapse onto an equal key) every offset is visited.
constexpr int kFillPiChunk = 4096, kFillPiThreads = 256;

__global__ void __launch_bounds__(kFillPiThreads) k_replay_fill(const FillParams p) {
  __shared__ float sSim[kFillPiChunk];
  __shared__ unsigned marks[kFillPiChunk / 32];
  __shared__ float sLow;
  __shared__ int sHs;
  const int job = blockIdx.x;
  if (job >= p.numJobs) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const int f = p.files[job];
  const int64_t fs = p.fileStart[f];
  const int64_t nValid = (p.fileStart[f + 1] - fs) - p.tailExtra - p.W + 1;
  const float *sim = p.sim + fs;

  Machine mc;
  mc.reset(p.entries + (size_t)job * (p.numPerFile + 1), p.numPerFile, p.maxEntrySz, 0, 0.f, p.minSpacing, p.step);
  // replicated view of the state every lane of warp 0 needs for its predicate
  bool hs = mc.has_space();
  float low = mc.lowest();
  int hasLast = 0;
  float lastSim = 0.f;
  int lastPi = 0;
  if (threadIdx.x == 0) { sLow = low; sHs = hs ? 1 : 0; }
  for (int64_t chunk = 0; chunk < nValid; chunk += kFillPiChunk) {
    __syncthreads();
    const float lowC = sLow;
    const bool hsC = sHs != 0;
    const int n = (int)(nValid - chunk < kFillPiChunk ? nValid - chunk : kFillPiChunk);
#pragma unroll
    for (int j = 0; j < kFillPiChunk / kFillPiThreads; j++) {
      const int i = j * kFillPiThreads + threadIdx.x;
      const float s = i < n ? sim[chunk + i] : 0.f;
      sSim[i] = s;
      const unsigned m = __ballot_sync(full, i < n && (hsC || s > lowC));
      if (lane == 0) marks[i >> 5] = m;
    }
    __syncthreads();
    if (warp != 0) continue;
    bool marksValid = !hs;
    for (int g = 0; g < (n + 31) / 32; g++) {
      if (marksValid && marks[g] == 0u) continue;
      const int i = 32 * g + lane;
      const int64_t t = chunk + i;
      const float s = sSim[i];
      unsigned todo = full;                        // offsets of the group not yet passed by the replay
      for (;;) {
        const bool accept = hs || s > low;
        const bool collapse = hasLast && ((t - (int64_t)lastPi - p.W) * p.step < p.minSpacing);
        const bool change = i < n && accept && (collapse ? (lastSim < s) : true);
        const unsigned cm = __ballot_sync(full, change) & todo;
        if (cm == 0u) break;
        const int hit = __ffs(cm) - 1;
        const float ss = __shfl_sync(full, s, hit);
        const int64_t ts = chunk + 32 * g + hit;
        if (lane == 0) {
          EntryRec m{ss, (int32_t)ts, (int32_t)(ts + p.W), 0.f, 1.0f};   // boost: filled in for the survivors
          mc.add(m);
          hs = mc.has_space();
          low = mc.lowest();
          hasLast = mc.hasLast;
          lastSim = mc.last.sim;
          lastPi = mc.last.piOff;
        }
        hs = __shfl_sync(full, (int)hs, 0) != 0;
        low = __shfl_sync(full, low, 0);
        hasLast = __shfl_sync(full, hasLast, 0);
        lastSim = __shfl_sync(full, lastSim, 0);
        lastPi = __shfl_sync(full, lastPi, 0);
        if (hs) marksValid = false;                // `low` may fall again: the marks are no superset any more
        todo = hit == 31 ? 0u : (full << (hit + 1));
        if (todo == 0u) break;
      }
    }
    if (lane == 0) { sLow = low; sHs = hs ? 1 : 0; }
  }
  if (threadIdx.x == 0) p.counts[job] = mc.n;
  if (warp == 0) {          // addMatch does not look at the boosts: computed for the entries that are left
    EntryRec *e = p.entries + (size_t)job * (p.numPerFile + 1);
    const int n = __shfl_sync(full, mc.n, 0);      // the machine lives in lane 0
    __syncwarp();
    for (int i = lane; i < n; i += 32) e[i].boostIn = p.boost.at(fs + e[i].piOff, e[i].piOff);
  }
}

// ---------------------------------------------------------------------------------------------
// full phase, punch-in only: emit every offset with sim > threshold[file] (unordered; the host
// sorts the few records by (file, offset) before the replay)
// ---------------------------------------------------------------------------------------------
struct CandParams {
  const float *sim;
  BoostSrc boost;
  const int64_t *fileStart;
  const int32_t *files;       // local file indices [numJobs]
  const float *thresholds;    // [numJobs]
  int numJobs;
  int W;
  int tailExtra;
  int fileBase;               // global index of local file 0
  sgz_record *out;
  int cap;
  int *counter;
};

__global__ void k_candidates(const CandParams p) {
  const int job = blockIdx.x;
  const int f = p.files[job];
  const float thr = p.thresholds[job];
  const int64_t fs = p.fileStart[f];
  const int64_t nValid = (p.fileStart[f + 1] - fs) - p.tailExtra - p.W + 1;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (int64_t base = (int64_t)blockIdx.y * blockDim.x; base < nValid; base += (int64_t)gridDim.y * blockDim.x) {
    const int64_t t = base + threadIdx.x;
    const float s = t < nValid ? p.sim[fs + t] : 0.f;
    const bool hit = t < nValid && s > thr;
    const unsigned mask = __ballot_sync(full, hit);
    if (mask == 0u) continue;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(p.counter, __popc(mask));
    slot0 = __shfl_sync(full, slot0, 0);
    if (hit) {
      const int slot = slot0 + __popc(mask & ((1u << lane) - 1u));
      if (slot < p.cap) {
        sgz_record r;
        r.file = p.fileBase + f;
        r.kind = 0;
        r.piOff = (int32_t)t;
        r.poOff = -1;
        r.sim = s;
        r.boostIn = p.boost.at(fs + t, t);
        r.boostOut = 1.0f;
        r.aux = 0;
        p.out[slot] = r;
      }
    }
  }
}

}  // namespace sgz
