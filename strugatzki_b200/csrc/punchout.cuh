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
// Filling rounds (allPrio not full): one warp replays the whole grid of a file, skipping 32 rows per
// step while the row gate fails and 32 cells per step while no cell can change the machine state.
// Full rounds: the GPU emits (a) every t with simIn[t] > theta and (b) every cell of a gated row with
// sim > theta, theta = allPrio.last.sim at the start of the round (a provable lower bound of every
// later lowestSim); the host replays the few records in (file, pi, po) order.
#pragma once
#include "common.cuh"
#include "corr.cuh"
#include "select.cuh"

namespace sgz {

struct FillPoParams {
  const float *simIn, *boostIn, *simOut, *boostOut, *rowMax;
  const int64_t *fileStart;
  const int32_t *files;
  int numJobs;
  int Win, Wout, minPunchF, maxPunchF;
  int numPerFile, maxEntrySz;
  int64_t minSpacing;
  int step;
  EntryRec *entries;   // [numJobs][numPerFile + 1]
  int32_t *counts;
};

__host__ __device__ __forceinline__ int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }

__device__ __forceinline__ float cell_sim(float inSim, float outSim) {
  return (float)sqrt((double)__fmul_rn(inSim, outSim));   // math.sqrt(inSim * outSim).toFloat, Float product
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

__global__ void k_replay_fill_po(const FillPoParams p) {
  const int warpsPerBlock = blockDim.x >> 5;
  const int job = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (job >= p.numJobs) return;
  const unsigned full = 0xffffffffu;
  const int f = p.files[job];
  const int64_t fs = p.fileStart[f];
  const int64_t N = p.fileStart[f + 1] - fs;
  const int64_t nA = N - p.minPunchF - p.Win + 1;
  const int64_t poMax = N - p.Wout;                       // last frame with an out-correlation
  const int span = p.maxPunchF - p.minPunchF + 1;
  Machine mc;
  mc.reset(p.entries + (size_t)job * (p.numPerFile + 1), p.numPerFile, p.maxEntrySz, 0, 0.f, p.minSpacing, p.step);
  // filling round: entryHasSpace holds at the file start, so tInOff = 0; loop B needs one full window
  const bool any = nA > 0 && (N - p.minPunchF) >= p.Wout && span > 0 && p.maxEntrySz > 0;
  bool hs = mc.has_space();
  float low = mc.lowest();
  int hasLast = 0;
  float lastSim = 0.f;
  int lastStop = 0;
  int64_t pi = 0;
  while (any && pi < nA) {
    // ---- skip rows whose gate fails (state is constant while nothing is accepted) ----
    const int64_t r = pi + lane;
    const float in = r < nA ? p.simIn[fs + r] : 0.f;
    const int64_t cellsR = i64min(poMax - (r + p.minPunchF) + 1, (int64_t)span);
    bool gate = r < nA && cellsR > 0 && in > __fmul_rn(low, low);
    // exact pruning: without space in the entry a row only matters if its best cell beats `low`
    if (gate && !hs) gate = cell_sim(in, p.rowMax[fs + r]) > low;
    const unsigned rmask = __ballot_sync(full, gate);
    if (rmask == 0u) { pi += 32; continue; }
    const int rl = __ffs(rmask) - 1;
    const int64_t row = pi + rl;
    const float inS = __shfl_sync(full, in, rl);
    const int n = (int)__shfl_sync(full, (int)cellsR, rl);
    const float bIn = p.boostIn[fs + row];
    int c = 0;
    while (c < n) {
      const int k = c + lane;
      const bool act = k < n;
      const int64_t po = row + p.minPunchF + k;
      const float s = act ? cell_sim(inS, p.simOut[fs + po]) : 0.f;
      bool change = false;
      if (act) {
        const bool accept = hs || s > low;
        const bool collapse = hasLast && ((row - (int64_t)lastStop) * p.step < p.minSpacing);
        change = accept && (collapse ? (lastSim < s) : true);
      }
      const unsigned cmask = __ballot_sync(full, change);
      if (cmask == 0u) { c += 32; continue; }
      const int cl = __ffs(cmask) - 1;
      const float ss = __shfl_sync(full, s, cl);
      const int64_t pos = row + p.minPunchF + c + cl;
      if (lane == 0) {
        EntryRec m{ss, (int32_t)row, (int32_t)pos, bIn, p.boostOut[fs + pos]};
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
      c = c + cl + 1;
    }
    pi = row + 1;
  }
  if (lane == 0) p.counts[job] = mc.n;
}

struct CandPoParams {
  const float *simIn, *boostIn, *simOut, *boostOut, *rowMax;
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

// one block per file; warps stride over rows
__global__ void k_candidates_po(const CandPoParams p) {
  const int job = blockIdx.x;
  const int f = p.files[job];
  const float thr = p.thresholds[job];
  const float thr2 = __fmul_rn(thr, thr);
  const int64_t fs = p.fileStart[f];
  const int64_t N = p.fileStart[f + 1] - fs;
  const int64_t nA = N - p.minPunchF - p.Win + 1;
  const int64_t poMax = N - p.Wout;
  const int span = p.maxPunchF - p.minPunchF + 1;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  if (nA <= 0 || span <= 0) return;
  // (a) in-curve offsets above the threshold: decide tInOff in the replay
  for (int64_t base = (int64_t)warp * 32; base < nA; base += (int64_t)nWarps * 32) {
    const int64_t t = base + lane;
    const float s = t < nA ? p.simIn[fs + t] : 0.f;
    const bool hit = t < nA && s > thr;
    const unsigned mask = __ballot_sync(full, hit);
    if (mask == 0u) continue;
    sgz_record r{p.fileBase + f, 3, (int32_t)t, -1, s, 0.f, 0.f, 0};
    emit_record(p, mask, lane, hit, r);
  }
  // (b) cells of gated rows
  for (int64_t rbase = (int64_t)warp * 32; rbase < nA; rbase += (int64_t)nWarps * 32) {
    const int64_t r = rbase + lane;
    const float in = r < nA ? p.simIn[fs + r] : 0.f;
    unsigned rmask = __ballot_sync(full, r < nA && in > thr2 && cell_sim(in, p.rowMax[fs + r]) > thr);
    while (rmask) {
      const int rl = __ffs(rmask) - 1;
      rmask &= rmask - 1;
      const int64_t row = rbase + rl;
      const float inS = __shfl_sync(full, in, rl);
      const int n = (int)i64min(poMax - (row + p.minPunchF) + 1, (int64_t)span);
      for (int c = 0; c < n; c += 32) {
        const int k = c + lane;
        const int64_t po = row + p.minPunchF + k;
        const float s = k < n ? cell_sim(inS, p.simOut[fs + po]) : 0.f;
        const bool hit = k < n && s > thr;
        const unsigned mask = __ballot_sync(full, hit);
        if (mask == 0u) continue;
        sgz_record rec{p.fileBase + f, 0, (int32_t)row, (int32_t)po, s, hit ? p.boostIn[fs + row] : 0.f,
                       hit ? p.boostOut[fs + po] : 0.f, __float_as_int(inS)};
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
      fp.simIn = job->simIn.p; fp.boostIn = job->boostIn.p; fp.simOut = job->simOut.p; fp.boostOut = job->boostOut.p;
      fp.rowMax = job->rowMaxOut.p;
      fp.fileStart = db->dFileStart.p; fp.files = job->dFiles.p; fp.numJobs = nj;
      fp.Win = job->qin.W; fp.Wout = job->qout.W; fp.minPunchF = job->minPunchF; fp.maxPunchF = job->maxPunchF;
      fp.numPerFile = npf; fp.maxEntrySz = m; fp.minSpacing = job->cfg.minSpacing; fp.step = job->step;
      fp.entries = job->dEntries.p; fp.counts = job->dCounts.p;
      SGZ_TRY(ctx->begin_call());
      k_replay_fill_po<<<(unsigned)ceil_div(nj, 4), 128, 0, ctx->stream>>>(fp);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_TRY(ctx->end_call());
      job->selectMs += ctx->lastMs;
      std::vector<int32_t> counts(nj);
      std::vector<EntryRec> ents((size_t)nj * (npf + 1));
      SGZ_CUDA(cudaMemcpyAsync(counts.data(), job->dCounts.p, nj * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(ents.data(), job->dEntries.p, ents.size() * sizeof(EntryRec), cudaMemcpyDeviceToHost,
                               ctx->stream));
      SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int j = 0; j < nj; j++)
        for (int k = 0; k < counts[j]; k++) {
          const EntryRec &e = ents[(size_t)j * (npf + 1) + k];
          job->localRecords.push_back(sgz_record{myLo + files[j], 1, e.piOff, e.stopOff, e.sim, e.boostIn, e.boostOut, 0});
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
        cp.simIn = job->simIn.p; cp.boostIn = job->boostIn.p; cp.simOut = job->simOut.p; cp.boostOut = job->boostOut.p;
        cp.rowMax = job->rowMaxOut.p;
        cp.fileStart = db->dFileStart.p; cp.files = job->dFiles.p; cp.thresholds = job->dThr.p; cp.numJobs = nj;
        cp.Win = job->qin.W; cp.Wout = job->qout.W; cp.minPunchF = job->minPunchF; cp.maxPunchF = job->maxPunchF;
        cp.fileBase = myLo; cp.out = job->dRecs.p; cp.cap = cap; cp.counter = job->dCounter.p;
        SGZ_TRY(ctx->begin_call());
        k_candidates_po<<<(unsigned)nj, 256, 0, ctx->stream>>>(cp);
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
      for (; i < recs.size() && recs[i].file == g; i++)
        if (recs[i].kind == 1) ent.push_back(EntryRec{recs[i].sim, recs[i].piOff, recs[i].poOff, recs[i].boostIn, recs[i].boostOut});
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
