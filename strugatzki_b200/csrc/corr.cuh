// corr.cuh -- host side of one FeatureCorrelation search: query preparation
// (readInBuffer, FeatureCorrelationImpl.scala:83-98), K1 launches, and the round-based
// select/merge protocol that reproduces allPrio / entryPrio exactly (see select.cuh).
#pragma once
#include "common.cuh"
#include "corr_kernel.cuh"
#include "corr_tc.cuh"
#include "corr_tc2.cuh"
#include "corr_refine.cuh"
#include "db.cuh"
#include "select.cuh"

struct PunchQuery {
  int W = 0, Wq = 0;
  float weight = 0.5f;
  std::vector<float> taps;  // [numCh][Wq]
  double stdT = 0, stdS = 0, rhoT = 0, rhoS = 0, lnAvg = 0;
  double rhoTFast = 0, rhoSFast = 0;   // sums of the taps rounded to FP16 (filter mode of k_corr_tc2: first parts only)
  DevBuf<float> dTaps;
  std::vector<uint16_t> tcTaps;
  DevBuf<uint16_t> dTcTaps;    // tensor-core path: hi/lo Toeplitz atoms (corr_tc.cuh), empty when not applicable
  DevBuf<unsigned char> dT2Taps;   // N = 64 tensor-core path (corr_tc2.cuh), empty when not applicable
  DevBuf<float> dA;                // normalised query [numCh][W]
  DevBuf<double> dAc;              // centred query in Double: a[c][i] + (-group mean), MathUtil.correlate's first factor, for
                                   // the exact re-evaluation of ill-conditioned windows (corr_fix.cuh)
};

struct sgz_corr {
  sgz_db *db = nullptr;
  sgz_ctx *ctx = nullptr;
  sgz_corr_config cfg{};
  int step = 1, minPunchF = 0, maxPunchF = 0;
  bool hasOut = false;
  PunchQuery qin, qout;
  int ntg = 128;
  int nslot = 3;
  bool useTc = false;       // K1 on the tensor cores (corr_tc.cuh) for resident scans
  bool useT2 = false;       // K1 on the tensor cores, N = 64 tiles fed by bulk copies (corr_tc2.cuh): the default
  int64_t numTilesTc = 0, numTilesT2 = 0;
  DevBuf<int32_t> dTileFile;   // [numTilesTc + 1] file holding the first frame of each tensor-core tile
  DevBuf<int32_t> dTileFileT2; // the same for the 8192-offset tiles of corr_tc2.cuh
  DevBuf<uint32_t> dFixList[2], dFixCount;   // ill-conditioned offsets of the punch-in / punch-out scan (corr_fix.cuh)
  // exact re-evaluation of the decisive offsets of a punch-in search (corr_refine.cuh)
  DevBuf<uint32_t> dRefList, dRefCount;
  DevBuf<int32_t> dRefCand;
  DevBuf<float> dRefThr;
  DevBuf<unsigned long long> dFileMaxExact;
  bool refine = false;      // this scan re-evaluates them
  bool fast = false;        // SGZ_FAST=1: K1 in filter mode (first-part product only), the re-evaluation with wider margins
  float tailMs = 0.f;       // device time of a scan behind its K1 launches (re-evaluations, per-file boosts)
  int64_t numTiles = 0;
  int64_t numOffsets = 0;
  DevBuf<float> simIn, boostIn, simOut, boostOut, rowMaxOut;
  DevBuf<unsigned long long> dFileMax, dFileMaxOut;
  // numPerFile = 1, punch-in only, tensor-core scan: the entry of a file is its maximum, which K1 leaves in dFileMax with
  // its first position; the scan adds the boost of that offset and a flag for files that hold a NaN window (whose entry
  // the reference decides by its NaN-first rule: those files keep the replay kernel).  Host copies from local_summary.
  DevBuf<uint32_t> dFileNaN;
  DevBuf<float> dFileBoost;
  std::vector<unsigned long long> hKeys;
  std::vector<uint32_t> hNaN;
  std::vector<float> hBoost;
  bool direct = false;      // this scan prepared the three arrays
  bool keysCached = false;  // ... and local_summary has fetched them
  bool summaryPrefetched = false;   // the pinned scratch holds the per-file results of the last scan
  bool scanned = false;

  // global (all ranks) view
  std::vector<sgz_file_summary> globalSummary;
  std::vector<sgz_file_summary> localSummary;   // cached by sgz_corr_local_top
  int nFilesGlobal = 0, myFirst = 0;
  bool globalSet = false;

  // replicated selection state
  std::vector<sgz_match> allPrio;  // descending Float.compare order, unique sims
  int nextFile = 0;
  bool finished = false;
  // current round
  int roundKind = -1;  // 0 = filling, 1 = full
  int roundFirst = 0, roundCount = 0, roundMaxEntrySz = 0;
  std::vector<sgz_record> localRecords;

  // device scratch for the selection kernels
  DevBuf<int32_t> dFiles, dCounts;
  DevBuf<float> dThr;
  DevBuf<sgz::EntryRec> dEntries;
  DevBuf<float4> dMeta;        // punch-out filling rounds: gate interval per file (punchout.cuh)
  DevBuf<sgz_record> dRecs;
  DevBuf<int> dCounter;

  // pinned host scratch for the small downloads of the selection rounds (pageable copies are staged synchronously)
  unsigned char *hPin = nullptr;
  size_t hPinBytes = 0;
  int pin(size_t bytes) {
    if (bytes <= hPinBytes) return SGZ_OK;
    if (hPin) cudaFreeHost(hPin);
    hPin = nullptr;
    hPinBytes = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 2, (size_t)1 << 18);
    SGZ_CUDA(cudaHostAlloc((void **)&hPin, want, cudaHostAllocDefault));
    hPinBytes = want;
    return SGZ_OK;
  }
  ~sgz_corr() { if (hPin) cudaFreeHost(hPin); }

  // timing (device ms of the last scan / accumulated select kernels)
  float scanMs = 0.f, selectMs = 0.f;
  int64_t scanLaunches = 0;

  // async
  std::thread worker;
  std::atomic<int> abortFlag{0}, doneFlag{0}, status{0};
  std::atomic<float> progress{0.f};
};

namespace sgz {

// the tensor-core K1 handles up to 14 channels (16 TMEM accumulators) and windows whose operands fit shared memory
inline bool tc_applicable(const sgz_ctx *ctx, int numCh, int W) {
  return numCh >= 2 && numCh <= 14 && W >= 1 && W <= 256 && tc_geom(W).smemBytes <= ctx->smemOptin;
}

// the N = 64 kernel takes any channel count (the spectral channels share five accumulators) and any window: beyond 257
// frames it runs in passes of 16 K steps (T2Geom); the limit is the zero slack behind the database (kDbSlack) that the
// tiles of the last file read into
inline bool t2_applicable(const sgz_ctx *ctx, int numCh, int W) {
  // (32-bit element indices into the per-frame arrays: databases beyond 2^32 - 2^16 frames = 13 000 h per GPU are refused
  // by db_ensure_planes)
  return numCh >= 2 && W >= 1 && W <= 7680 && t2_geom(W, ctx->smemOptin).smemBytes <= ctx->smemOptin;
}

// FP16 planes and per-frame sums of the frames [db->planesUpto, upto) -- whole 2048-frame blocks; upto < 0 = everything
// a tile may touch (the zero slack behind the last file included)
inline int db_ensure_planes(sgz_db *db, int64_t upto, cudaStream_t st) {
  const int64_t strideFrames = (db->capFrames + kT2Tile - 1) / kT2Tile * kT2Tile;
  SGZ_REQUIRE(strideFrames < ((int64_t)1 << 32) - 65536, "database of %lld frames exceeds the 32-bit frame index of the tensor-core scan",
              (long long)db->usedFrames);
  if (!db->dPlanes.p || db->planeStrideBytes != strideFrames * 2) {
    db->planeStrideBytes = strideFrames * 2;
    SGZ_TRY(db->dPlanes.alloc((size_t)db->numCh * 2 * (size_t)db->planeStrideBytes));
    SGZ_TRY(db->dSideA.alloc((size_t)strideFrames));
    SGZ_TRY(db->dSideB.alloc((size_t)strideFrames));
    db->planeRows = strideFrames / kT2P;
    SGZ_TRY(db->dB16.alloc((size_t)16 * db->planeRows));
    db->planesUpto = 0;
  }
  const int64_t all = std::min(strideFrames, (db->usedFrames + kDbSlack + kPlaneFrames - 1) / kPlaneFrames * kPlaneFrames);
  const int64_t end = upto < 0 ? all : std::min(all, upto / kPlaneFrames * kPlaneFrames);
  if (end <= db->planesUpto) return SGZ_OK;
  const unsigned blocks = (unsigned)((end - db->planesUpto) / kPlaneFrames);
  k_db_planes<<<blocks, 256, 0, st>>>(db->dData.p, db->capFrames, db->numCh, db->numPairs, db->planesUpto,
                                      std::min(end, db->capFrames), db->dPlanes.p, db->planeStrideBytes, db->dSideA.p,
                                      db->dSideB.p, db->dB16.p, db->planeRows);
  SGZ_LAUNCH_CHECK(db->ctx);
  db->planesUpto = end;
  return SGZ_OK;
}

// readInBuffer (FeatureCorrelationImpl.scala:83-98): cut [start,stop) feature frames, normalise,
// matrix-wide stats of the temporal (ch 0) and spectral (ch 1..) groups, ln of the loudness average.
inline int prepare_query(const sgz_db *db, const float *inputPlanar /*[numCh][inputFrames]*/, int64_t inputFrames,
                         int64_t spanStart, int64_t spanStop, float weight, int step, PunchQuery &q, cudaStream_t st) {
  const int numCh = db->numCh;
  const int start = full_to_feat(spanStart, step), stop = full_to_feat(spanStop, step);
  const int W = stop - start;
  SGZ_REQUIRE(W > 0, "punch span [%lld,%lld) is empty after rounding to feature frames", (long long)spanStart,
              (long long)spanStop);
  if (start < 0 || stop > inputFrames) {
    set_error("punch span reads feature frames [%d,%d) but the input has %lld frames (reference: EOFException)",
              start, stop, (long long)inputFrames);
    return SGZ_ERR_IO;
  }
  q.W = W;
  q.Wq = (W + 3) / 4 * 4;
  q.weight = weight;
  std::vector<float> a((size_t)numCh * W);
  for (int c = 0; c < numCh; c++) {
    float mn = 0.f, d = 1.f;
    if (db->hasNorm) { mn = db->norm[2 * c]; d = db->norm[2 * c + 1] - mn; }
    for (int i = 0; i < W; i++) {
      float f = inputPlanar[(size_t)c * inputFrames + start + i];
      a[(size_t)c * W + i] = db->hasNorm ? (f - mn) / d : f;   // MathUtil.normalize
    }
  }
  auto stat = [&](int c0, int c1, double &mean, double &sd) {  // MathUtil.stat, same loop order
    double sum = 0.0;
    for (int c = c0; c < c1; c++) for (int i = 0; i < W; i++) sum += a[(size_t)c * W + i];
    int matSize = W * (c1 - c0);
    mean = sum / matSize;
    sum = 0.0;
    for (int c = c0; c < c1; c++) for (int i = 0; i < W; i++) { double dd = a[(size_t)c * W + i] - mean; sum += dd * dd; }
    sd = sqrt(sum / matSize);
  };
  double meanT, meanS;
  stat(0, 1, meanT, q.stdT);
  stat(1, numCh, meanS, q.stdS);
  {
    double sum = 0.0;  // MathUtil.avg -> Float, then math.log
    for (int i = 0; i < W; i++) sum += a[i];
    float avg = (float)(sum / W);
    q.lnAvg = log((double)avg);
  }
  // taps in the kernel's pair layout: [pair p][tap i] = float2(channel 2p, channel 2p+1), zero padded
  q.taps.assign((size_t)db->numPairs * q.Wq * 2, 0.f);
  q.rhoT = q.rhoS = q.rhoTFast = q.rhoSFast = 0.0;
  for (int c = 0; c < numCh; c++) {
    double mean = c == 0 ? meanT : meanS;
    for (int i = 0; i < W; i++) {
      float tp = (float)((double)a[(size_t)c * W + i] - mean);
      q.taps[((size_t)(c >> 1) * q.Wq + i) * 2 + (c & 1)] = tp;
      (c == 0 ? q.rhoT : q.rhoS) += (double)tp;
      (c == 0 ? q.rhoTFast : q.rhoSFast) += (double)__half2float(__float2half_rn(tp));
    }
  }
  SGZ_TRY(q.dTaps.alloc(q.taps.size()));
  SGZ_CUDA(cudaMemcpyAsync(q.dTaps.p, q.taps.data(), q.taps.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  // not while uploads are in flight (streaming scan, FFMA2 kernel): a 200 KB pageable copy would queue behind them in
  // the H2D copy engine and stall the job creation until the database has landed
  if (tc_applicable(db->ctx, numCh, W) && db->chunks.empty()) {
    tc_build_taps(q.taps, db->numPairs, q.Wq, W, q.tcTaps);
    SGZ_TRY(q.dTcTaps.alloc(q.tcTaps.size()));
    SGZ_CUDA(cudaMemcpyAsync(q.dTcTaps.p, q.tcTaps.data(), q.tcTaps.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
  }
  if (t2_applicable(db->ctx, numCh, W)) {
    SGZ_TRY(q.dA.alloc(a.size()));
    SGZ_TRY(q.dAc.alloc(a.size()));
    SGZ_CUDA(cudaMemcpyAsync(q.dA.p, a.data(), a.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    k_t2_query<<<ceil_div(numCh * W, 256), 256, 0, st>>>(q.dA.p, numCh, W, -meanT, -meanS, q.dAc.p);
    SGZ_LAUNCH_CHECK(db->ctx);
    const T2Geom g = t2_geom(W, db->ctx->smemOptin);
    SGZ_TRY(q.dT2Taps.alloc((size_t)numCh * g.tapsFullBytes));
    k_t2_taps<<<numCh, 256, 0, st>>>(reinterpret_cast<const float2 *>(q.dTaps.p), numCh, q.Wq, W, q.dT2Taps.p);
    SGZ_LAUNCH_CHECK(db->ctx);
  }
  return SGZ_OK;
}

// consumer threads per CTA: the largest configuration whose ring + stats buffers fit in shared memory
inline int pick_ntg(const sgz_ctx *ctx, int numPairs, int Wq) {
  // preferred: 128 consumers per CTA and TWO persistent CTAs per SM -- the CTAs drift out of phase, so one
  // CTA's epilogue / row prologue / stats wait overlaps the other's FFMA2 loop
  const size_t perSm = 228 * 1024, reserved = 1024;
  if (const char *e = getenv("SGZ_CORR_NC")) {   // tuning override (developer knob)
    const int nc = atoi(e);
    if (nc >= 32 && nc <= 320 && nc % 32 == 0 && corr_smem_layout(nc, numPairs, Wq).total <= ctx->smemOptin) return nc;
  }
  // measured on B200 (600 x 51 680 frames, W = 172): 256 x 1 CTA/SM 10.2e9 offsets/s, 128 x 2 CTAs/SM 9.8e9,
  // 192 x 1 8.7e9, 128 x 1 9.0e9, 96 x 2 7.6e9
  if (corr_smem_layout(256, numPairs, Wq).total <= ctx->smemOptin) return 256;
  if (2 * (corr_smem_layout(128, numPairs, Wq).total + reserved) <= perSm) return 128;
  const int opts[5] = {256, 192, 128, 64, 32};
  for (int k = 0; k < 5; k++)
    if (corr_smem_layout(opts[k], numPairs, Wq).total <= ctx->smemOptin) return opts[k];
  return 0;
}

inline int corr_ctas_per_sm(int ntg, int numPairs, int Wq) {
  const size_t perSm = 228 * 1024, reserved = 1024;
  if (const char *e = getenv("SGZ_CORR_CTAS")) return atoi(e) == 2 ? 2 : 1;
  return (ntg <= 128 && 2 * (corr_smem_layout(ntg, numPairs, Wq).total + reserved) <= perSm) ? 2 : 1;
}

inline int64_t valid_offsets(const sgz_db *db, int W, int tailExtra) {
  int64_t total = 0;
  for (int f = 0; f < db->numFiles(); f++) {
    int64_t n = (db->fileStart[f + 1] - db->fileStart[f]) - tailExtra - W + 1;
    if (n > 0) total += n;
  }
  return total;
}

// one K1 launch over tiles [tileBegin, tileEnd) on `st`; `spareSMs` SMs are left to other streams (the prepare
// kernels of uploads still in flight during a streaming scan)
inline int run_scan_one(sgz_corr *job, PunchQuery &q, int tailExtra, float *sim, float *boost,
                        unsigned long long *fileMax, int64_t tileBegin, int64_t tileEnd, cudaStream_t st,
                        int spareSMs) {
  if (tileEnd <= tileBegin) return SGZ_OK;
  sgz_db *db = job->db;
  CorrParams p{};
  p.data = db->dData.p;
  p.rowStride = db->capFrames;
  p.usedFrames = db->usedFrames;
  p.numCh = db->numCh;
  p.numPairs = db->numPairs;
  p.W = q.W;
  p.Wq = q.Wq;
  p.taps = reinterpret_cast<const float2 *>(q.dTaps.p);
  p.stdT = q.stdT; p.stdS = q.stdS; p.rhoT = q.rhoT; p.rhoS = q.rhoS;
  p.lnAvgIn = q.lnAvg;
  p.weight = q.weight;
  p.maxBoost = job->cfg.maxBoost;
  p.fileStart = db->dFileStart.p;
  p.numFiles = db->numFiles();
  p.tailExtra = tailExtra;
  p.sim = sim;
  p.boost = boost;
  p.fileMax = fileMax;
  p.tileBegin = tileBegin;
  p.tileEnd = tileEnd;
  CorrSmemLayout L = corr_smem_layout(job->ntg, db->numPairs, q.Wq, job->nslot);
  sgz_ctx *ctx = job->ctx;
  auto kern = job->nslot == 2 ? k_corr<2> : k_corr<3>;
  SGZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smemOptin));
  // persistent: one or two CTAs per SM, tiles are striped over the CTAs
  const int perSm = corr_ctas_per_sm(job->ntg, db->numPairs, q.Wq);
  const int sms = std::max(ctx->smCount - spareSMs, 1);
  const unsigned grid = (unsigned)std::min<int64_t>(tileEnd - tileBegin, (int64_t)sms * perSm);
  kern<<<grid, job->ntg + 64, L.total, st>>>(p);
  SGZ_LAUNCH_CHECK(ctx);
  return SGZ_OK;
}

// tensor-core variant of run_scan_one over the whole database (one fused kernel, see corr_tc.cuh)
inline int run_scan_tc(sgz_corr *job, PunchQuery &q, int tailExtra, float *sim, float *boost,
                       unsigned long long *fileMax, cudaStream_t st) {
  sgz_db *db = job->db;
  sgz_ctx *ctx = job->ctx;
  const TcGeom G = tc_geom(q.W);
  CorrTcParams tp{};
  tp.data = db->dData.p; tp.rowStride = db->capFrames; tp.usedFrames = db->usedFrames;
  tp.numCh = db->numCh; tp.numPairs = db->numPairs; tp.W = q.W;
  tp.taps = q.dTcTaps.p;
  tp.stdT = q.stdT; tp.stdS = q.stdS; tp.rhoT = q.rhoT; tp.rhoS = q.rhoS; tp.lnAvgIn = q.lnAvg;
  tp.weight = q.weight; tp.maxBoost = job->cfg.maxBoost;
  tp.fileStart = db->dFileStart.p; tp.tileFile = job->dTileFile.p; tp.numFiles = db->numFiles(); tp.tailExtra = tailExtra;
  tp.tileBegin = 0; tp.tileEnd = job->numTilesTc;
  tp.sim = sim; tp.boost = boost; tp.fileMax = fileMax;
  SGZ_CUDA(cudaFuncSetAttribute(k_corr_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smemBytes));
  const unsigned gridTc = (unsigned)std::min<int64_t>(job->numTilesTc, ctx->smCount);
  DevBuf<long long> dProf;
  static const bool chainOrder = getenv("SGZ_CORR_TC_CHAIN") && atoi(getenv("SGZ_CORR_TC_CHAIN")) == 1;
  tp.chainOrder = chainOrder;
  const bool prof = getenv("SGZ_CORR_TC_PROF") != nullptr;   // developer probe: where does the issuer lane wait?
  if (prof) {
    SGZ_TRY(dProf.alloc((size_t)gridTc * 24));
    SGZ_CUDA(cudaMemsetAsync(dProf.p, 0, (size_t)gridTc * 24 * sizeof(long long), st));
    tp.prof = dProf.p;
  }
  k_corr_tc<<<gridTc, kTcThreads, G.smemBytes, st>>>(tp);
  SGZ_LAUNCH_CHECK(ctx);
  if (prof) {
    std::vector<long long> h((size_t)gridTc * 24);
    SGZ_CUDA(cudaMemcpyAsync(h.data(), dProf.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    SGZ_CUDA(cudaStreamSynchronize(st));
    double a[24] = {0};
    for (unsigned b = 0; b < gridTc; b++) for (int k = 0; k < 24; k++) a[k] += (double)h[(size_t)b * 24 + k];
    const double tiles = a[6] > 0 ? a[6] : 1;
    fprintf(stderr, "k_corr_tc cycles per tile: issuer total %.0f | wait opFree %.0f, accEmpty %.0f, opFull %.0f, taps %.0f, "
                    "issue %.0f || epilogue wait accFull %.0f, tmem read %.0f, wait stats %.0f, row sums %.0f, init %.0f, loop %.0f || split loads+sums %.0f, wait opFree %.0f, store %.0f, stats %.0f\n",
            a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles, a[4] / tiles, a[5] / tiles, a[8] / tiles, a[9] / tiles,
            a[10] / tiles, a[12] / tiles, a[13] / tiles, a[11] / tiles, a[16] / tiles, a[17] / tiles, a[18] / tiles, a[19] / tiles);
  }
  return SGZ_OK;
}

// N = 64 tensor-core K1 over tiles [tileBegin, tileEnd) of 8192 offsets (corr_tc2.cuh); the planes of every frame the
// tiles touch must have been enqueued on `st` (db_ensure_planes)
// where the boost of an offset comes from: the curve a round-1 kernel wrote, or (tensor-core scan: arr = nullptr) the
// loudness channel of the database
inline BoostSrc boost_src(const sgz_corr *job, const PunchQuery &q, const float *arr) {
  return BoostSrc{arr, job->db->dData.p, q.W, q.lnAvg};
}

constexpr uint32_t kFixCap = 1u << 20;

// exact Double replay of the offsets the tensor-core scan flagged as ill-conditioned (corr_fix.cuh); `which` = 0 punch-in,
// 1 punch-out curve.  Enqueued behind the K1 launches of that curve; costs one tiny launch when nothing was flagged.
inline int run_fixup(sgz_corr *job, PunchQuery &q, int which, int tailExtra, float *sim, unsigned long long *fileMax,
                     cudaStream_t st) {
  sgz_db *db = job->db;
  CorrFixParams fp{};
  fp.data = db->dData.p; fp.rowStride = db->capFrames; fp.usedFrames = db->usedFrames;
  fp.a = q.dAc.p; fp.numCh = db->numCh; fp.W = q.W; fp.stdT = q.stdT; fp.stdS = q.stdS;
  fp.weight = q.weight; fp.maxBoost = job->cfg.maxBoost;
  fp.fileStart = db->dFileStart.p; fp.numFiles = db->numFiles(); fp.tailExtra = tailExtra;
  fp.list = job->dFixList[which].p; fp.count = job->dFixCount.p + which; fp.cap = kFixCap;
  fp.sim = sim; fp.boost = boost_src(job, q, nullptr); fp.fileMax = fileMax;
  fp.fileNaN = (which == 0 && job->direct) ? job->dFileNaN.p : nullptr;
  k_corr_fixup<<<(unsigned)job->ctx->smCount * 4, 128, 0, st>>>(fp);
  SGZ_LAUNCH_CHECK(job->ctx);
  return SGZ_OK;
}

// exact sims for every offset that can reach the result of a punch-in search, file maxima rebuilt from them
// (corr_refine.cuh); enqueued behind run_fixup, before anything reads the curve or the maxima
inline int run_refine(sgz_corr *job, cudaStream_t st) {
  sgz_db *db = job->db;
  const int nf = db->numFiles();
  if (nf == 0 || db->usedFrames <= 0) return SGZ_OK;
  PunchQuery &q = job->qin;
  SGZ_CUDA(cudaMemsetAsync(job->dRefCount.p, 0, 2 * sizeof(uint32_t), st));     // [offsets listed, files listed]
  SGZ_CUDA(cudaMemsetAsync(job->dFileMaxExact.p, 0, (size_t)nf * sizeof(unsigned long long), st));
  // filter mode: FP16 signal and taps -- a sim is off by about 1.4e-4 x (level / spread) / sqrt(cells of the window)
  float margin = kRefineMargin, tieTol = kRefineTieTol;
  if (job->fast) {
    const float err = std::min(0.05f, 2.5e-4f * sqrtf(172.f / (float)q.W));
    margin = 4.f * err;
    tieTol = 2.f * err;
  }
  k_refine_threshold<<<1, kRefineThrThreads, 0, st>>>(job->dFileMax.p, nf, job->cfg.numMatches, margin, tieTol,
                                         job->dRefThr.p, job->dRefCand.p, job->dRefCount.p + 1);
  SGZ_LAUNCH_CHECK(job->ctx);
  k_refine_collect<<<(unsigned)job->ctx->smCount * 8, 256, 0, st>>>(
      job->simIn.p, db->dFileStart.p, job->dFileMax.p, q.W, 0, job->dRefThr.p, margin, job->cfg.numPerFile == 1 ? 1 : 0,
      job->dRefCand.p, job->dRefCount.p + 1, job->dRefList.p, job->dRefCount.p, kRefineCap);
  SGZ_LAUNCH_CHECK(job->ctx);
  CorrFixParams fp{};
  fp.data = db->dData.p; fp.rowStride = db->capFrames; fp.usedFrames = db->usedFrames;
  fp.a = q.dAc.p; fp.numCh = db->numCh; fp.W = q.W; fp.stdT = q.stdT; fp.stdS = q.stdS;
  fp.weight = q.weight; fp.maxBoost = job->cfg.maxBoost;
  fp.fileStart = db->dFileStart.p; fp.numFiles = nf; fp.tailExtra = 0;
  fp.list = job->dRefList.p; fp.count = job->dRefCount.p; fp.cap = kRefineCap;
  fp.sim = job->simIn.p; fp.boost = boost_src(job, q, nullptr); fp.fileMax = job->dFileMaxExact.p;
  fp.fileNaN = nullptr; fp.listOnly = 1;
  // a warp per offset with the window in shared memory (k_corr_exact); windows too long for that: a thread per offset
  const size_t perWarp = kExactBytesPerValue * (size_t)db->numCh * (size_t)q.W;
  const int warps = (int)std::min<size_t>(4, ((size_t)200 << 10) / perWarp);
  if (warps >= 1) {
    const size_t smem = (size_t)warps * perWarp;
    if (smem > ((size_t)48 << 10))
      SGZ_CUDA(cudaFuncSetAttribute(k_corr_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_corr_exact<<<(unsigned)job->ctx->smCount * 8, 32 * warps, smem, st>>>(fp);
  } else {
    k_corr_fixup<<<(unsigned)job->ctx->smCount * 4, 128, 0, st>>>(fp);
  }
  SGZ_LAUNCH_CHECK(job->ctx);
  k_refine_merge<<<ceil_div(nf, 256), 256, 0, st>>>(job->dFileMaxExact.p, job->dFileMax.p, nf);
  SGZ_LAUNCH_CHECK(job->ctx);
  if (getenv("SGZ_REFINE_DEBUG")) {   // developer probe: how much was re-evaluated
    uint32_t h[2];
    float t;
    SGZ_CUDA(cudaMemcpyAsync(h, job->dRefCount.p, sizeof h, cudaMemcpyDeviceToHost, st));
    SGZ_CUDA(cudaMemcpyAsync(&t, job->dRefThr.p, sizeof t, cudaMemcpyDeviceToHost, st));
    SGZ_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "refine: threshold %.7g, %u files listed, %u offsets re-evaluated%s\n", t, h[1], h[0],
            h[0] > kRefineCap ? " (list overflowed: none)" : "");
  }
  return SGZ_OK;
}

// boost of every file's best offset (numPerFile = 1 searches take their entries from the file maxima)
inline int run_filemax_boost(sgz_corr *job, cudaStream_t st) {
  sgz_db *db = job->db;
  const int nf = db->numFiles();
  if (nf == 0) return SGZ_OK;
  const int warps = (size_t)job->qin.W * 16 <= ((size_t)48 << 10) ? 4 : 1;
  k_filemax_boost_warp<<<ceil_div(nf, warps), 32 * warps, (size_t)warps * job->qin.W * sizeof(float), st>>>(
      boost_src(job, job->qin, nullptr), db->dFileStart.p, job->dFileMax.p, nf, job->dFileBoost.p);
  SGZ_LAUNCH_CHECK(job->ctx);
  return SGZ_OK;
}

// boost curves of both punch windows of a punch-out search on the tensor-core scan (k_boost_all, corr_fix.cuh)
inline int run_boost_curves(sgz_corr *job, cudaStream_t st) {
  sgz_db *db = job->db;
  if (db->usedFrames <= 0) return SGZ_OK;
  const unsigned blocks = (unsigned)ceil_div<int64_t>(ceil_div<int64_t>(db->usedFrames, kBoostRun), 128);
  k_boost_all<<<blocks, 128, 0, st>>>(db->dData.p, db->usedFrames, job->qin.W, job->qin.lnAvg, job->boostIn.p);
  SGZ_LAUNCH_CHECK(job->ctx);
  k_boost_all<<<blocks, 128, 0, st>>>(db->dData.p, db->usedFrames, job->qout.W, job->qout.lnAvg, job->boostOut.p);
  SGZ_LAUNCH_CHECK(job->ctx);
  return SGZ_OK;
}

inline int run_scan_t2(sgz_corr *job, PunchQuery &q, int which, int tailExtra, float *sim,
                       unsigned long long *fileMax, int64_t tileBegin, int64_t tileEnd, cudaStream_t st, int spareSMs) {
  if (tileEnd <= tileBegin) return SGZ_OK;
  sgz_db *db = job->db;
  sgz_ctx *ctx = job->ctx;
  static const int ring = getenv("SGZ_T2_RING") ? (int)strtol(getenv("SGZ_T2_RING"), nullptr, 16) : 0x43;   // developer knob
  const T2Geom G = t2_geom(q.W, ctx->smemOptin, ring);
  CorrT2Params tp{};
  tp.ring = ring;
  tp.tapsFirst = getenv("SGZ_T2_TAPS_FIRST") ? atoi(getenv("SGZ_T2_TAPS_FIRST")) : 0;
  tp.planes = db->dPlanes.p; tp.planeStrideBytes = db->planeStrideBytes;
  tp.sideA = db->dSideA.p; tp.sideB = db->dSideB.p;
  tp.b16 = db->dB16.p; tp.rowsTotal = db->planeRows;
  tp.usedFrames = db->usedFrames; tp.numCh = db->numCh; tp.W = q.W;
  tp.taps = q.dT2Taps.p;
  {
    T2Eval &E = tp.ev;
    E.nT = (double)q.W; E.nS = (double)(db->numCh - 1) * (double)q.W;
    E.invNTd = 1.0 / E.nT; E.invNSd = 1.0 / E.nS;
    E.nTf = (float)E.nT; E.nSf = (float)E.nS;
    {
      const double eT = std::min(0.5, std::max(2e-3, 0.25 / E.nT)), eS = std::min(0.5, std::max(2e-3, 0.25 / E.nS));
      E.sqEpsT = (float)sqrt(eT / (1.0 - eT)); E.sqEpsS = (float)sqrt(eS / (1.0 - eS));
    }
    // boost = exp((lnAvgIn - ln avg) / 0.6) <= maxBoost  <=>  avg >= exp(lnAvgIn - 0.6 ln maxBoost); a NaN threshold
    // (maxBoost < 0 or NaN) lets no offset pass, like the reference's comparison
    const double avgMin = exp(q.lnAvg - 0.6 * log((double)job->cfg.maxBoost));
    E.gateNm = (float)(avgMin * E.nT);
    E.invNT2 = (float)(1.0 / (E.nT * E.nT)); E.invNS2 = (float)(1.0 / (E.nS * E.nS));
    E.cT = (float)(1.0 / (E.nT * q.stdT)); E.cS = (float)(1.0 / (E.nS * q.stdS));
    const bool fastRho = job->fast && which == 0 && G.NP == 1;
    E.kTn = (float)((fastRho ? q.rhoTFast : q.rhoT) / (E.nT * E.nT * q.stdT));
    E.kSn = (float)((fastRho ? q.rhoSFast : q.rhoS) / (E.nS * E.nS * q.stdS));
    E.wT = q.weight; E.wS = 1.0f - q.weight;
    E.useT = q.weight > 0.f; E.useS = q.weight < 1.f;
  }
  tp.fileStart = db->dFileStart.p; tp.tileFile = job->dTileFileT2.p; tp.numFiles = db->numFiles(); tp.tailExtra = tailExtra;
  tp.tileBegin = tileBegin; tp.tileEnd = tileEnd;
  tp.sim = sim; tp.fileMax = fileMax;
  tp.fixList = job->dFixList[which].p; tp.fixCount = job->dFixCount.p + which; tp.fixCap = kFixCap;
  tp.smemMax = (int)ctx->smemOptin;
  {
    static const int ahead = getenv("SGZ_T2_AHEAD") ? atoi(getenv("SGZ_T2_AHEAD")) : kT2Ahead;   // developer knob
    tp.ahead = ahead;
    tp.dbg = getenv("SGZ_T2_DBG") ? atoi(getenv("SGZ_T2_DBG")) : 0;
    static const int split = getenv("SGZ_T2_SPLIT") ? atoi(getenv("SGZ_T2_SPLIT")) : 1;                 // developer knob
    tp.splitRelease = split;
    static const int l2hint = getenv("SGZ_T2_L2HINT") ? atoi(getenv("SGZ_T2_L2HINT")) : 7;             // developer knob
    tp.l2hint = l2hint;
    static const int narrow = getenv("SGZ_T2_NARROW") ? atoi(getenv("SGZ_T2_NARROW")) : 1;               // developer knob
    tp.narrow = narrow;
  }
  const bool prof = getenv("SGZ_CORR_TC_PROF") != nullptr;   // developer probe: cycles per role and phase
  auto kern = G.NP > 1 ? (prof ? k_corr_tc2<true, true> : k_corr_tc2<false, true>) : (prof ? k_corr_tc2<true, false> : k_corr_tc2<false, false>);
  if (job->fast && which == 0 && G.NP == 1) kern = prof ? k_corr_tc2<true, false, true> : k_corr_tc2<false, false, true>;
  SGZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smemBytes));
  const unsigned grid = (unsigned)std::min<int64_t>(tileEnd - tileBegin, std::max(ctx->smCount - spareSMs, 1));
  DevBuf<long long> dProf;
  if (prof) {
    SGZ_TRY(dProf.alloc((size_t)grid * 24));
    SGZ_CUDA(cudaMemsetAsync(dProf.p, 0, (size_t)grid * 24 * sizeof(long long), st));
    tp.prof = dProf.p;
  }
  kern<<<grid, kT2Threads, G.smemBytes, st>>>(tp);
  SGZ_LAUNCH_CHECK(ctx);
  if (prof) {
    std::vector<long long> h((size_t)grid * 24);
    SGZ_CUDA(cudaMemcpyAsync(h.data(), dProf.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    SGZ_CUDA(cudaStreamSynchronize(st));
    double a[24] = {0};
    for (unsigned b = 0; b < grid; b++) for (int k = 0; k < 24; k++) a[k] += (double)h[(size_t)b * 24 + k];
    const double tiles = a[5] > 0 ? a[5] : 1;
    fprintf(stderr, "k_corr_tc2 cycles per tile (8192 offsets): issuer total %.0f | wait accEmpty %.0f, signal %.0f, taps %.0f, issue %.0f "
                    "|| epilogue singles + wait sums %.0f, window init %.0f, wait accFull %.0f, tmem read %.0f, 16 offsets %.0f\n",
            a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles, a[4] / tiles, a[8] / tiles, a[9] / tiles, a[10] / tiles,
            a[11] / tiles, a[12] / tiles);
    // per CTA: a static round-robin of tiles ends with the slowest SM
    std::vector<double> cyc, ns;
    for (unsigned b = 0; b < grid; b++)
      if (h[(size_t)b * 24 + 5] > 0) {
        cyc.push_back((double)h[(size_t)b * 24] / (double)h[(size_t)b * 24 + 5]);
        ns.push_back((double)h[(size_t)b * 24 + 6]);
      }
    if (!cyc.empty()) {
      std::vector<double> c2 = cyc, n2 = ns;
      std::sort(c2.begin(), c2.end()); std::sort(n2.begin(), n2.end());
      fprintf(stderr, "k_corr_tc2 per CTA: cycles per tile min %.0f, median %.0f, max %.0f | issuer wall ns min %.0f, median %.0f, max %.0f | mean clock %.3f GHz\n",
              c2.front(), c2[c2.size() / 2], c2.back(), n2.front(), n2[n2.size() / 2], n2.back(), a[0] / std::max(a[6], 1.0));
      if (getenv("SGZ_T2_PROF_CTAS")) {
        for (size_t b = 0; b < cyc.size(); b++) fprintf(stderr, "%s%.0f", b ? " " : "k_corr_tc2 cycles per tile by CTA: ", cyc[b]);
        fprintf(stderr, "\n");
      }
    }
  }
  return SGZ_OK;
}

// ---- allPrio helpers (host): SortedSet[Match](MatchMinOrd) ----
inline int allprio_find(const std::vector<sgz_match> &s, float sim, bool &found) {
  found = false;
  size_t i = 0;
  for (; i < s.size(); i++) {
    int c = jfloat_compare(s[i].sim, sim);
    if (c == 0) { found = true; return (int)i; }
    if (c < 0) return (int)i;
  }
  return (int)i;
}
inline void allprio_add(std::vector<sgz_match> &s, const sgz_match &m) {
  bool found;
  int i = allprio_find(s, m.sim, found);
  if (!found) s.insert(s.begin() + i, m);
}

}  // namespace sgz
