// api_segself.cuh -- extern "C" entry points of FeatureSegmentation and SelfSimilarity
// (included by api.cu; unity build).
#pragma once
#include "common.cuh"
#include "db.cuh"
#include "segm.cuh"
#include "selfsim.cuh"
#include "selfsim_fast.cuh"
#include "selfsim_tc.cuh"
#include "cross.cuh"

namespace sgz {

// raw frames (any layout) [first, first+n) -> normalised planar device buffer [numCh][stride]
inline int upload_features(sgz_ctx *ctx, int numCh, const float *norm, const void *frames, int64_t nFrames,
                           int layout, int64_t first, int64_t n, DevBuf<float> &out, int64_t &stride) {
  SGZ_REQUIRE(numCh >= 2 && numCh <= 48, "numCh = numCoeffs + 1 must be in [2,48], got %d", numCh);
  SGZ_REQUIRE(layout >= 0 && layout <= 2, "unknown layout %d", layout);
  SGZ_REQUIRE(first >= 0 && n >= 0 && first + n <= nFrames, "frame range outside the file");
  stride = std::max<int64_t>((n + 3) / 4 * 4, 4);
  SGZ_TRY(out.alloc((size_t)stride * numCh));
  SGZ_CUDA(cudaMemsetAsync(out.p, 0, (size_t)stride * numCh * sizeof(float), ctx->stream));
  if (n == 0) return SGZ_OK;
  std::vector<float> nm((size_t)numCh * 2);
  for (int c = 0; c < numCh; c++) { nm[2 * c] = norm ? norm[2 * c] : 0.f; nm[2 * c + 1] = norm ? norm[2 * c + 1] : 1.f; }
  DevBuf<float> dNorm, stage;
  SGZ_TRY(dNorm.alloc(nm.size()));
  SGZ_CUDA(cudaMemcpyAsync(dNorm.p, nm.data(), nm.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  SGZ_TRY(stage.alloc((size_t)n * numCh));
  const float *src = (const float *)frames;
  if (layout == SGZ_LAYOUT_PLANAR_LE) {
    SGZ_CUDA(cudaMemcpy2DAsync(stage.p, (size_t)n * sizeof(float), src + first, (size_t)nFrames * sizeof(float),
                               (size_t)n * sizeof(float), (size_t)numCh, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    SGZ_CUDA(cudaMemcpyAsync(stage.p, src + first * numCh, (size_t)n * numCh * sizeof(float), cudaMemcpyHostToDevice,
                             ctx->stream));
  }
  int blocks = (int)ceil_div<int64_t>(n, kPrepFrames);
  k_db_prepare<<<blocks, 256, (size_t)kPrepFrames * numCh * sizeof(float), ctx->stream>>>(
      stage.p, layout, n, out.p, stride, 0, numCh, dNorm.p, 0);
  SGZ_LAUNCH_CHECK(ctx);
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));  // stage / dNorm go out of scope
  return SGZ_OK;
}

inline void span_to_feat(int hasStart, int hasStop, int64_t spanStart, int64_t spanStop, int step, int64_t nFrames,
                         int &afStart, int &afStop) {
  afStart = 0;
  if (hasStart) afStart = std::max(0, full_to_feat(spanStart, step));
  afStop = (int)nFrames;
  if (hasStop) afStop = std::min((int)nFrames, full_to_feat(spanStop, step));
}

inline int self_geometry(const sgz_self_config *cfg, int64_t n1, int64_t n2, sgz_self_geometry *g, int *Hout) {
  SGZ_REQUIRE(cfg->stepSize > 0, "stepSize must be > 0");
  const int H = full_to_feat(cfg->corrLen, cfg->stepSize);
  SGZ_REQUIRE(H > 0, "corrLen rounds to an empty window");
  SGZ_REQUIRE(cfg->decimation >= 1, "Illegal decimation setting of %d", cfg->decimation);
  const int64_t afNum = std::min(n1, n2);
  int afStart, afStop;
  span_to_feat(cfg->hasStart, cfg->hasStop, cfg->spanStart, cfg->spanStop, cfg->stepSize, afNum, afStart, afStop);
  const int afLen = afStop - afStart;
  const int64_t n = std::max<int64_t>(0, (int64_t)afLen - 2 * H + 1);
  const int numCorrs = (int)n;
  int d = cfg->decimation, ext = numCorrs / d;
  if (ext > 0xB504) {                                   // SelfSimilarityImpl.scala:85-90
    d = (numCorrs + 0xB503) / 0xB504;
    ext = numCorrs / d;
  }
  g->imgExt = ext;
  g->decim = d;
  g->numCorrs = numCorrs;
  g->afStart = afStart;
  g->numCells = (int64_t)ext * (ext + 1) / 2;
  if (Hout) *Hout = H;
  return SGZ_OK;
}

}  // namespace sgz

extern "C" {

int sgz_segm_run(sgz_ctx *ctx, const sgz_segm_config *cfg, int32_t numCh, const float *norm, const void *frames,
                 int64_t nFrames, int32_t layout, sgz_break *out, int32_t cap, int32_t *n, float *curve,
                 int64_t curveCap, int64_t *numOffsets) {
  using namespace sgz;
  SGZ_REQUIRE(ctx && cfg && frames && n, "sgz_segm_run: NULL argument");
  SGZ_REQUIRE(cfg->stepSize > 0, "stepSize must be > 0");
  SGZ_TRY(ctx->bind());
  const int step = cfg->stepSize;
  const int H = full_to_feat(cfg->corrLen, step);
  SGZ_REQUIRE(H > 0, "corrLen rounds to an empty window");
  int afStart, afStop;
  span_to_feat(cfg->hasStart, cfg->hasStop, cfg->spanStart, cfg->spanStop, step, nFrames, afStart, afStop);
  const int afLen = afStop - afStart;
  *n = 0;
  if (numOffsets) *numOffsets = 0;
  if (afLen <= 0) return SGZ_OK;
  // reference loop count (:107-133): one iteration for the first (possibly short) read, then one per frame
  const int nOff = afLen >= 2 * H ? afLen - 2 * H + 1 : 1;
  DevBuf<float> x, dCurve;
  int64_t stride = 0;
  SGZ_TRY(upload_features(ctx, numCh, norm, frames, nFrames, layout, afStart, afLen, x, stride));
  SGZ_TRY(dCurve.alloc(nOff));
  DevBuf<sgz_break> dOut;
  DevBuf<int> dCount;
  const int nb = std::max(cfg->numBreaks, 0);
  SGZ_TRY(dOut.alloc((size_t)nb + 1));
  SGZ_TRY(dCount.alloc(1));
  SegmParams sp{x.p, stride, afLen, numCh, H, nOff, cfg->temporalWeight, dCurve.p};
  const int pickShared = ((size_t)nb + 1) * sizeof(sgz_break) <= 8 * 1024;
  const size_t pickSmem = kPickChunk * sizeof(float) + (pickShared ? ((size_t)nb + 1) * sizeof(sgz_break) : 0);
  PickParams pp{dCurve.p, nOff, afStart, H, step, nb, cfg->minSpacing, dOut.p, dCount.p, pickShared};
  SGZ_TRY(ctx->begin_call());
  k_segm_curve<<<ceil_div(nOff, 128), 128, 0, ctx->stream>>>(sp);
  SGZ_LAUNCH_CHECK(ctx);
  static const bool pickSmemOnly = getenv("SGZ_SEGM_PICK_SMEM") != nullptr;   // developer knob: the shared-memory replay
  if (nb <= 31 && !pickSmemOnly) k_segm_pick_warp<<<1, 32, kPickChunk * sizeof(float), ctx->stream>>>(pp);
  else k_segm_pick<<<1, 32, pickSmem, ctx->stream>>>(pp);
  SGZ_LAUNCH_CHECK(ctx);
  SGZ_TRY(ctx->end_call());
  int count = 0;
  SGZ_CUDA(cudaMemcpyAsync(&count, dCount.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<sgz_break> br((size_t)std::max(count, 1));
  if (count > 0) {
    SGZ_CUDA(cudaMemcpyAsync(br.data(), dOut.p, count * sizeof(sgz_break), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (curve && curveCap > 0) {
    SGZ_CUDA(cudaMemcpyAsync(curve, dCurve.p, (size_t)std::min<int64_t>(curveCap, nOff) * sizeof(float),
                             cudaMemcpyDeviceToHost, ctx->stream));
  }
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  *n = count;
  if (numOffsets) *numOffsets = nOff;
  if (out) {
    SGZ_REQUIRE(cap >= count, "break buffer too small (%d < %d)", cap, count);
    for (int i = 0; i < count; i++) out[i] = br[i];
  }
  return SGZ_OK;
}

int sgz_self_geometry_of(const sgz_self_config *cfg, int64_t nFrames1, int64_t nFrames2, sgz_self_geometry *out) {
  using namespace sgz;
  SGZ_REQUIRE(cfg && out, "NULL argument");
  return self_geometry(cfg, nFrames1, nFrames2, out, nullptr);
}

static int self_prepare(sgz_ctx *ctx, const sgz_self_config *cfg, int32_t numCh, const float *norm,
                        const void *frames1, int64_t n1, const void *frames2, int64_t n2, int32_t layout,
                        sgz_self_geometry &g, int &H, DevBuf<float> &x1, DevBuf<float> &x2, DevBuf<int32_t> &dLut,
                        sgz::SelfParams &p) {
  using namespace sgz;
  SGZ_REQUIRE(cfg->colorWarp > 0, "Illegal color warp setting. Must be > 0, but is %g", cfg->colorWarp);
  SGZ_REQUIRE(cfg->colorCeil > 0, "Illegal color ceil setting. Must be > 0, but is %g", cfg->colorCeil);
  if (!frames2) n2 = n1;
  SGZ_TRY(self_geometry(cfg, n1, n2, &g, &H));
  // frames needed: [afStart, afStart + numCorrs - 1 + H] of both files (rightOff + H <= afLen - H)
  const int64_t need = g.numCorrs > 0 ? (int64_t)g.numCorrs - 1 + H : 0;
  int64_t s1 = 0, s2 = 0;
  SGZ_TRY(upload_features(ctx, numCh, norm, frames1, n1, layout, g.afStart, need, x1, s1));
  if (frames2) SGZ_TRY(upload_features(ctx, numCh, norm, frames2, n2, layout, g.afStart, need, x2, s2));
  if (cfg->lut) {
    SGZ_REQUIRE(cfg->lutSize >= 2, "lutSize must be >= 2");
    SGZ_TRY(dLut.alloc(cfg->lutSize));
    SGZ_CUDA(cudaMemcpyAsync(dLut.p, cfg->lut, cfg->lutSize * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  p = SelfParams{};
  p.x1 = x1.p;
  p.x2 = frames2 ? x2.p : x1.p;
  p.stride1 = s1;
  p.stride2 = frames2 ? s2 : s1;
  p.numCh = numCh; p.H = H; p.decim = g.decim; p.imgExt = g.imgExt;
  p.weight = cfg->temporalWeight;
  p.colorInv = cfg->colorInv;
  p.colorWarp = cfg->colorWarp;
  p.colorScale = 1.0f / cfg->colorCeil;
  p.lut = cfg->lut ? dLut.p : nullptr;
  p.lutSize = cfg->lutSize;
  return SGZ_OK;
}

// Default (non-precise) image path on p (p.rgb set by the caller): centred Gram tiles + closed-form epilogue.
// Tensor-core kernel (selfsim_tc.cuh) when the geometry fits and the data are finite, else the FFMA2 kernel
// (selfsim_fast.cuh).  simMat (device, [ext][ext], tensor-core kernel only) receives the raw sims of the upper triangle;
// *usedTc reports which kernel ran.
static int self_fast_image(sgz_ctx *ctx, const sgz_self_config *cfg, sgz::SelfParams &p, const sgz_self_geometry &g,
                           int H, int numCh, bool cross, int rowBegin, int rowEnd, float *simMat, int *usedTc) {
  using namespace sgz;
  const int ext = g.imgExt;
  const int64_t need = (int64_t)g.numCorrs - 1 + H;
  if (usedTc) *usedTc = 0;
  DevBuf<double> dMeans;
  DevBuf<unsigned int> dAmax;
  DevBuf<float2> ws1, ws2;
  DevBuf<float4> wsA, wsB;
  DevBuf<uint4> rec1, rec2, tail2;
  DevBuf<int2> dTiles;
  SGZ_TRY(dMeans.alloc(2));
  SGZ_TRY(dAmax.alloc(2));
  SGZ_CUDA(cudaMemsetAsync(dMeans.p, 0, 2 * sizeof(double), ctx->stream));
  SGZ_CUDA(cudaMemsetAsync(dAmax.p, 0, 2 * sizeof(unsigned int), ctx->stream));
  SGZ_TRY(ctx->begin_call());
  k_self_means<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(p.x1, p.stride1, need, numCh, dMeans.p);
  SGZ_LAUNCH_CHECK(ctx);
  static const bool tcOff = getenv("SGZ_SELF_TC") && atoi(getenv("SGZ_SELF_TC")) == 0;
  static const bool aDescOff = getenv("SGZ_SELF_TC_ADESC") && atoi(getenv("SGZ_SELF_TC_ADESC")) == 0;
  const SelfTcGeom G = self_tc_geom(H, g.decim, ctx->smemOptin, !aDescOff);
  // The tensor core truncates when it adds a product block to the accumulator, a bias that grows with the number of MMAs
  // into one (large) accumulator (tools/selfsim_error_probe.py, worst deviation from the oracle): with ONE spectral main
  // accumulator 78 MMAs (13 channels x 6 K steps, H = 86) give 3.3e-6 absolute / 3.8e-6 relative, H = 96 8.6e-6 relative,
  // 104 MMAs (H = 128) 1.0e-5, 416 (H = 512) 5.7e-5.  So from 60 MMAs on the main products alternate between TWO TMEM
  // regions by channel parity (twoMain: half the chain per accumulator -- 1.8e-6 absolute at H = 86, 3.8e-6 at H = 176 --
  // at the price of the spare region the temporal MMAs of the next tile would otherwise use: -2 % speed), and windows whose
  // half chain still exceeds 78 MMAs (H > 176 at 14 channels) go to the FFMA2 kernel, whose round-to-nearest errors average
  // out (8e-7 at H = 192).
  const int chainAll = (numCh - 1) * G.nks, chainHalf = (numCh / 2) * G.nks;   // numCh / 2 = odd spectral channels
  static const int oneMainMax = getenv("SGZ_SELF_TC_ONE_MAIN_MAX") ? atoi(getenv("SGZ_SELF_TC_ONE_MAIN_MAX")) : 60;
  const bool twoMain = chainAll > oneMainMax && numCh >= 3;
  const bool chainOk = (twoMain ? chainHalf : chainAll) <= 78 && 3 * G.nks <= 78;
  // Longer windows stay on the tensor cores in CHUNKS of 160 frames (10 K steps: 7 x 10 MMAs per main accumulator with
  // twoMain), one launch per chunk: a launch adds the raw Gram sums of the earlier chunks from a scratch image and the
  // last one applies the closed form of the whole window (SelfTcParams::recOff).  The scratch costs 16 B per cell and
  // chunk of HBM traffic next to about 3 000 tensor-core flops per cell and chunk.
  constexpr int kChunkH = 160;
  static const bool chunkOff = getenv("SGZ_SELF_TC_CHUNKS") && atoi(getenv("SGZ_SELF_TC_CHUNKS")) == 0;
  struct Chunk { int h0, Hc; SelfTcGeom G; bool twoMain; };
  std::vector<Chunk> chunks;
  if (chainOk && G.ok) chunks.push_back(Chunk{0, H, G, twoMain});
  else if (!chunkOff && numCh >= 2) {
    bool ok = true;
    for (int h0 = 0; h0 < H && ok; h0 += kChunkH) {
      const int Hc = std::min(kChunkH, H - h0);
      const SelfTcGeom Gc = self_tc_geom(Hc, g.decim, ctx->smemOptin, !aDescOff);
      const bool two = (numCh - 1) * Gc.nks > oneMainMax && numCh >= 3;
      ok = Gc.ok && (two ? (numCh / 2) * Gc.nks : (numCh - 1) * Gc.nks) <= 78;
      chunks.push_back(Chunk{h0, Hc, Gc, two});
    }
    if (!ok) chunks.clear();
  }
  bool tc = !tcOff && numCh >= 2 && !chunks.empty();
  if (tc) {
    k_self_absmax<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(p.x1, p.stride1, need, numCh, dAmax.p);
    SGZ_LAUNCH_CHECK(ctx);
    if (cross) {
      k_self_absmax<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(p.x2, p.stride2, need, numCh, dAmax.p);
      SGZ_LAUNCH_CHECK(ctx);
    }
  }
  double means[2];
  unsigned int amaxBits[2] = {0, 0};
  SGZ_CUDA(cudaMemcpyAsync(means, dMeans.p, sizeof means, cudaMemcpyDeviceToHost, ctx->stream));
  SGZ_CUDA(cudaMemcpyAsync(amaxBits, dAmax.p, sizeof amaxBits, cudaMemcpyDeviceToHost, ctx->stream));
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!std::isfinite(means[0]) || !std::isfinite(means[1])) {   // NaN / Inf features: not for the centred Gram forms
    SGZ_TRY(ctx->end_call());
    if (usedTc) *usedTc = -1;
    return SGZ_OK;
  }
  SelfFastParams fp{};
  fp.base = p;
  fp.shiftT = (float)(means[0] / (double)need);
  fp.shiftS = (float)(means[1] / ((double)need * (numCh - 1)));
  fp.cross = cross;
  std::vector<int2> gt;
  for (int a0 = rowBegin / kGT * kGT; a0 < rowEnd; a0 += kGT)
    for (int b0 = a0; b0 < ext; b0 += kGT) gt.push_back(make_int2(a0, b0));
  SGZ_TRY(dTiles.alloc(gt.size()));
  SGZ_CUDA(cudaMemcpyAsync(dTiles.p, gt.data(), gt.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
  float scale = 1.0f;
  if (tc) {
    // power-of-two scale that brings the centred data to |v| <= 8192 (FP16 range with headroom); Inf / NaN data or
    // shifts leave the tensor-core path
    float aT, aS;
    memcpy(&aT, &amaxBits[0], 4);
    memcpy(&aS, &amaxBits[1], 4);
    const double bT = (double)aT + std::fabs((double)fp.shiftT), bS = (double)aS + std::fabs((double)fp.shiftS);
    const double bound = std::max(bT, bS);
    if (!std::isfinite(bT) || !std::isfinite(bS)) tc = false;   // (std::max alone would drop a NaN operand)
    else if (bound > 0) scale = (float)std::ldexp(1.0, std::max(-60, std::min(60, (int)std::floor(std::log2(8192.0 / bound)))));
  }
  if (tc) {
    const int nT = ceil_div(ext, kGT);
    int64_t spanMax = 0;
    for (const Chunk &ck : chunks) spanMax = std::max<int64_t>(spanMax, ck.G.span + ck.h0 / G.g);
    const int64_t nRec = (int64_t)G.dp * kGT * (nT - 1) + spanMax;
    const int64_t recThreads = (int64_t)numCh * nRec;
    SGZ_TRY(rec1.alloc((size_t)2 * numCh * nRec));
    SGZ_TRY(wsA.alloc(ext));
    k_self_records<<<(unsigned)ceil_div<int64_t>(recThreads, 256), 256, 0, ctx->stream>>>(
        p.x1, p.stride1, need, numCh, G.g, nRec, fp.shiftT, fp.shiftS, scale, rec1.p);
    SGZ_LAUNCH_CHECK(ctx);
    k_self_wsums4<<<ceil_div(ext, 128), 128, 0, ctx->stream>>>(p.x1, p.stride1, numCh, H, g.decim, ext, fp.shiftT, fp.shiftS,
                                                              scale, wsA.p);
    SGZ_LAUNCH_CHECK(ctx);
    if (cross) {
      SGZ_TRY(rec2.alloc((size_t)2 * numCh * nRec));
      SGZ_TRY(wsB.alloc(ext));
      k_self_records<<<(unsigned)ceil_div<int64_t>(recThreads, 256), 256, 0, ctx->stream>>>(
          p.x2, p.stride2, need, numCh, G.g, nRec, fp.shiftT, fp.shiftS, scale, rec2.p);
      SGZ_LAUNCH_CHECK(ctx);
      k_self_wsums4<<<ceil_div(ext, 128), 128, 0, ctx->stream>>>(p.x2, p.stride2, numCh, H, g.decim, ext, fp.shiftT,
                                                                fp.shiftS, scale, wsB.p);
      SGZ_LAUNCH_CHECK(ctx);
    }
    SelfTcParams tp{};
    tp.f = fp;
    tp.rec1 = rec1.p;
    tp.rec2 = cross ? rec2.p : rec1.p;
    tp.nRec = nRec;
    tp.wsA = wsA.p;
    tp.wsB = cross ? wsB.p : wsA.p;
    tp.tiles = dTiles.p;
    tp.nTiles = (int)gt.size();
    tp.simMat = simMat;
    static const int dumpEnv = getenv("SGZ_SELF_TC_DUMP") ? atoi(getenv("SGZ_SELF_TC_DUMP")) : -1;
    const unsigned gridTc = (unsigned)std::min<size_t>(gt.size(), (size_t)ctx->smCount);
    DevBuf<long long> dProf;
    DevBuf<float> dCorrT, dGPart;
    const bool prof = getenv("SGZ_SELF_TC_PROF") != nullptr;   // developer probe: cycles per role
    // Both groups in play: one launch with three TMEM regions per tile (default; the epilogue of a tile is not overlapped
    // with the next tile's MMAs), or SGZ_SELF_TC_PASSES=2: temporal pass -> corrT, then the spectral pass with two tiles
    // in TMEM (see SelfTcParams).  Measured on B200 (30 000 frames, decim 1): 1.16e11 vs 1.06e11 cells/s -- the overlapped
    // epilogue runs at half speed next to the MMAs' shared-memory traffic and the extra pass costs 8 k cycles per tile.
    static const bool onePassEnv = !(getenv("SGZ_SELF_TC_PASSES") && atoi(getenv("SGZ_SELF_TC_PASSES")) == 2);
    const bool onePass = onePassEnv || chunks.size() > 1;
    const bool useT = p.weight > 0.f, useS = p.weight < 1.f;
    const int nPass = useT && useS && !onePass ? 2 : 1;
    if (nPass == 2) {
      SGZ_TRY(dCorrT.alloc((size_t)ext * ext));
      tp.corrT = dCorrT.p;
    }
    if (chunks.size() > 1) {
      SGZ_TRY(dGPart.alloc((size_t)2 * ext * ext));
      tp.gPart = dGPart.p;
      tp.Hfull = H;
    }
    for (size_t ci = 0; ci < chunks.size(); ci++) {
      const Chunk &ck = chunks[ci];
      const SelfTcGeom &Gc = ck.G;
      tp.f.base.H = ck.Hc;
      tp.recOff = ck.h0 / Gc.g;
      tp.storeG = ci + 1 < chunks.size();
      tp.loadG = ci > 0;
      tp.nks = Gc.nks; tp.nSlab = Gc.nSlab; tp.slabKs = Gc.slabKs; tp.dp = Gc.dp; tp.kcStep = Gc.kcStep; tp.span = Gc.span;
      tp.nStage = Gc.nStage; tp.nRecStage = Gc.nRecStage; tp.matBytes = Gc.matBytes; tp.stageBytes = Gc.stageBytes;
      tp.recPartBytes = Gc.recPartBytes; tp.recStageBytes = Gc.recStageBytes; tp.tailBytes = Gc.tailBytes;
      tp.tail2 = nullptr; tp.nTailRows = 0;
      if (Gc.tailBytes) {   // in-place mode, H % 16 != 0: B's last K step, cut off at the window's end, for every window row
        const int64_t nRows = (int64_t)nT * kGT;
        SGZ_TRY(tail2.alloc((size_t)2 * numCh * 2 * nRows));
        // (the rows of a chunk start recOff records into the window; the channel stride stays nRec)
        k_self_tail<<<(unsigned)ceil_div<int64_t>(2 * (int64_t)numCh * 2 * nRows, 256), 256, 0, ctx->stream>>>(
            tp.rec2, numCh, nRec, Gc.kcStep, 2 * (Gc.nks - 1), ck.Hc, nRows, tail2.p, tp.recOff);
        SGZ_LAUNCH_CHECK(ctx);
        tp.tail2 = tail2.p;
        tp.nTailRows = nRows;
      }
      tp.aDesc = Gc.aDesc;
      tp.twoMain = ck.twoMain;
      tp.dump = dumpEnv >= 0 ? std::min(dumpEnv, Gc.dump) : Gc.dump;
      SGZ_CUDA(cudaFuncSetAttribute(k_self_gram_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gc.smemBytes));
      SGZ_CUDA(cudaFuncSetAttribute(k_self_gram_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gc.smemBytes));
      SGZ_CUDA(cudaFuncSetAttribute(k_self_gram_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gc.smemBytes));
    for (int pass = 0; pass < nPass; pass++) {
      const bool doT = nPass == 2 ? pass == 0 : useT, doS = nPass == 2 ? pass == 1 : useS;
      tp.storeT = nPass == 2 && pass == 0;
      tp.loadT = nPass == 2 && pass == 1;
      if (prof) {
        SGZ_TRY(dProf.alloc((size_t)gridTc * 16));
        SGZ_CUDA(cudaMemsetAsync(dProf.p, 0, (size_t)gridTc * 16 * sizeof(long long), ctx->stream));
        tp.prof = dProf.p;
      }
      if (doT && doS) k_self_gram_tc<0><<<gridTc, kSgThreads, Gc.smemBytes, ctx->stream>>>(tp);
      else if (doT) k_self_gram_tc<1><<<gridTc, kSgThreads, Gc.smemBytes, ctx->stream>>>(tp);
      else k_self_gram_tc<2><<<gridTc, kSgThreads, Gc.smemBytes, ctx->stream>>>(tp);
      SGZ_LAUNCH_CHECK(ctx);
      if (prof) {
        std::vector<long long> h((size_t)gridTc * 16);
        SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
        SGZ_CUDA(cudaMemcpy(h.data(), dProf.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double a[16] = {0};
        for (unsigned bk = 0; bk < gridTc; bk++) for (int k = 0; k < 16; k++) a[k] += (double)h[(size_t)bk * 16 + k];
        const double tiles = a[5] > 0 ? a[5] : 1;
        fprintf(stderr, "k_self_gram_tc chunk %d/%d pass %d/%d cycles per tile (aDesc %d, %d stages of %d K steps, %d record stages, %d parked batches): issuer total %.0f | "
                        "records %.0f, wait accEmpty %.0f, wait full %.0f, issue %.0f || builder total %.0f | wait records %.0f, "
                        "wait empty %.0f, build %.0f || epilogue wait accFull %.0f, main %.0f\n",
                (int)ci + 1, (int)chunks.size(), pass + 1, nPass, tp.aDesc, tp.nStage, tp.slabKs, tp.nRecStage, tp.dump, a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles,
                a[4] / tiles, a[8] / tiles, a[9] / tiles, a[10] / tiles, a[11] / tiles, a[12] / tiles, a[13] / tiles);
      }
    }
    }
    SGZ_TRY(ctx->end_call());
    if (usedTc) *usedTc = 1;
    ctx->lastSelfKernel = 2;
    return SGZ_OK;
  }
  SGZ_TRY(ws1.alloc((size_t)2 * ext));
  k_self_wsums<<<ceil_div(ext, 128), 128, 0, ctx->stream>>>(p.x1, p.stride1, numCh, H, g.decim, ext, fp.shiftT, fp.shiftS,
                                                           ws1.p);
  SGZ_LAUNCH_CHECK(ctx);
  fp.ws1 = ws1.p;
  fp.ws2 = ws1.p;
  if (cross) {
    SGZ_TRY(ws2.alloc((size_t)2 * ext));
    k_self_wsums<<<ceil_div(ext, 128), 128, 0, ctx->stream>>>(p.x2, p.stride2, numCh, H, g.decim, ext, fp.shiftT,
                                                             fp.shiftS, ws2.p);
    SGZ_LAUNCH_CHECK(ctx);
    fp.ws2 = ws2.p;
  }
  k_self_gram<<<(unsigned)gt.size(), 256, 0, ctx->stream>>>(fp, dTiles.p);
  SGZ_LAUNCH_CHECK(ctx);
  SGZ_TRY(ctx->end_call());
  ctx->lastSelfKernel = 1;
  return SGZ_OK;
}

int sgz_self_run(sgz_ctx *ctx, const sgz_self_config *cfg, int32_t numCh, const float *norm, const void *frames1,
                 int64_t nFrames1, const void *frames2, int64_t nFrames2, int32_t layout, int32_t rowBegin,
                 int32_t rowEnd, int32_t *rgb, int64_t rgbCap, sgz_self_geometry *geom) {
  using namespace sgz;
  SGZ_REQUIRE(ctx && cfg && frames1, "sgz_self_run: NULL argument");
  SGZ_TRY(ctx->bind());
  sgz_self_geometry g{};
  int H = 0;
  DevBuf<float> x1, x2;
  DevBuf<int32_t> dLut, dRgb;
  DevBuf<int2> dTiles;
  SelfParams p;
  SGZ_TRY(self_prepare(ctx, cfg, numCh, norm, frames1, nFrames1, frames2, nFrames2, layout, g, H, x1, x2, dLut, p));
  if (geom) *geom = g;
  const int ext = g.imgExt;
  if (rowEnd <= 0 || rowEnd > ext) rowEnd = ext;
  if (rowBegin < 0) rowBegin = 0;
  if (ext == 0 || rowBegin >= rowEnd) return SGZ_OK;
  SGZ_REQUIRE(!rgb || rgbCap >= (int64_t)ext * ext, "rgb buffer too small for %d x %d pixels", ext, ext);
  SGZ_TRY(dRgb.alloc((size_t)ext * ext));
  SGZ_CUDA(cudaMemsetAsync(dRgb.p, 0, (size_t)ext * ext * sizeof(int32_t), ctx->stream));
  p.colBegin = rowBegin;
  p.colEnd = rowEnd;
  p.rgb = dRgb.p;
  // Very short windows (H < 16 frames) go to the exact path: the closed form cancels badly when a window's mean is far from
  // the file mean relative to its own spread, and the exact replay costs O(H) per cell anyway.
  const bool precise = cfg->precise || H < 16;
  if (!precise) {
    int used = 0;
    SGZ_TRY(self_fast_image(ctx, cfg, p, g, H, numCh, frames2 != nullptr, rowBegin, rowEnd, nullptr, &used));
    if (used >= 0) {
      if (rgb) {
        SGZ_CUDA(cudaMemcpyAsync(rgb, dRgb.p, (size_t)ext * ext * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
      }
      return SGZ_OK;
    }
    // used < 0: NaN / Inf in the features.  The Gram forms centre the data by the file mean, which would smear one bad
    // frame over the whole image; the exact replay keeps it local to the windows over that frame, like the reference.
  }
  std::vector<int2> tiles;
  for (int a0 = rowBegin / kSelfTile * kSelfTile; a0 < rowEnd; a0 += kSelfTile)
    for (int b0 = a0 / kSelfTile * kSelfTile; b0 < ext; b0 += kSelfTile) tiles.push_back(make_int2(a0, b0));
  SGZ_TRY(dTiles.alloc(tiles.size()));
  SGZ_CUDA(cudaMemcpyAsync(dTiles.p, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
  const size_t smem = (size_t)2 * numCh * (g.decim * (kSelfTile - 1) + H) * sizeof(float);
  SGZ_REQUIRE(smem <= ctx->smemOptin, "self-similarity tile needs %zu bytes of shared memory", smem);
  SGZ_CUDA(cudaFuncSetAttribute(k_self_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smemOptin));
  SGZ_TRY(ctx->begin_call());
  k_self_tiles<<<(unsigned)tiles.size(), kSelfTile * kSelfTile, smem, ctx->stream>>>(p, dTiles.p);
  SGZ_LAUNCH_CHECK(ctx);
  SGZ_TRY(ctx->end_call());
  ctx->lastSelfKernel = 3;
  if (rgb) {
    SGZ_CUDA(cudaMemcpyAsync(rgb, dRgb.p, (size_t)ext * ext * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SGZ_OK;
}

int sgz_self_cells(sgz_ctx *ctx, const sgz_self_config *cfg, int32_t numCh, const float *norm, const void *frames1,
                   int64_t nFrames1, const void *frames2, int64_t nFrames2, int32_t layout, int64_t nCells,
                   const int32_t *leftIdx, const int32_t *rightIdx, float *sim, int32_t *rgb) {
  using namespace sgz;
  SGZ_REQUIRE(ctx && cfg && frames1 && leftIdx && rightIdx, "sgz_self_cells: NULL argument");
  SGZ_TRY(ctx->bind());
  sgz_self_geometry g{};
  int H = 0;
  DevBuf<float> x1, x2, dSim;
  DevBuf<int32_t> dLut, dL, dR, dRgb;
  SelfParams p;
  SGZ_TRY(self_prepare(ctx, cfg, numCh, norm, frames1, nFrames1, frames2, nFrames2, layout, g, H, x1, x2, dLut, p));
  if (nCells <= 0) return SGZ_OK;
  for (int64_t k = 0; k < nCells; k++)
    SGZ_REQUIRE(leftIdx[k] >= 0 && leftIdx[k] < g.imgExt && rightIdx[k] >= 0 && rightIdx[k] < g.imgExt,
                "cell %lld = (%d,%d) outside the %d x %d image", (long long)k, leftIdx[k], rightIdx[k], g.imgExt,
                g.imgExt);
  SGZ_TRY(dL.alloc(nCells));
  SGZ_TRY(dR.alloc(nCells));
  SGZ_TRY(dSim.alloc(nCells));
  SGZ_TRY(dRgb.alloc(nCells));
  SGZ_CUDA(cudaMemcpyAsync(dL.p, leftIdx, nCells * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  SGZ_CUDA(cudaMemcpyAsync(dR.p, rightIdx, nCells * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  p.leftIdx = dL.p; p.rightIdx = dR.p; p.nCells = nCells; p.simOut = dSim.p; p.rgbOut = dRgb.p;
  SGZ_TRY(ctx->begin_call());
  if (cfg->precise || H < 16) {
    k_self_cells<<<(unsigned)ceil_div<int64_t>(nCells, 128), 128, 0, ctx->stream>>>(p);
    SGZ_LAUNCH_CHECK(ctx);
  } else {
    const int64_t need = (int64_t)g.numCorrs - 1 + H;
    DevBuf<double> dMeans;
    SGZ_TRY(dMeans.alloc(2));
    SGZ_CUDA(cudaMemsetAsync(dMeans.p, 0, 2 * sizeof(double), ctx->stream));
    k_self_means<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(p.x1, p.stride1, need, numCh, dMeans.p);
    SGZ_LAUNCH_CHECK(ctx);
    double means[2];
    SGZ_CUDA(cudaMemcpyAsync(means, dMeans.p, sizeof means, cudaMemcpyDeviceToHost, ctx->stream));
    SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!std::isfinite(means[0]) || !std::isfinite(means[1])) {   // NaN / Inf features: exact replay (see sgz_self_run)
      k_self_cells<<<(unsigned)ceil_div<int64_t>(nCells, 128), 128, 0, ctx->stream>>>(p);
      SGZ_LAUNCH_CHECK(ctx);
    } else {
      SelfFastParams fp{};
      fp.base = p;
      fp.shiftT = (float)(means[0] / (double)need);
      fp.shiftS = (float)(means[1] / ((double)need * (numCh - 1)));
      k_self_cells_fast<<<(unsigned)ceil_div<int64_t>(nCells, 128), 128, 0, ctx->stream>>>(fp);
      SGZ_LAUNCH_CHECK(ctx);
    }
  }
  SGZ_TRY(ctx->end_call());
  if (sim) SGZ_CUDA(cudaMemcpyAsync(sim, dSim.p, nCells * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (rgb) SGZ_CUDA(cudaMemcpyAsync(rgb, dRgb.p, nCells * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!cfg->precise && H >= 16 && g.imgExt <= 4096) {
    // Images of this geometry come from the tile kernel (tensor cores when eligible): report ITS values for every cell
    // of the computed triangle, so that the parity checks on cells exercise the kernel that renders the image.
    const int ext = g.imgExt;
    DevBuf<int32_t> dImg;
    DevBuf<float> dMat;
    SGZ_TRY(dImg.alloc((size_t)ext * ext));
    SGZ_TRY(dMat.alloc((size_t)ext * ext));
    p.colBegin = 0;
    p.colEnd = ext;
    p.rgb = dImg.p;
    int usedTc = 0;
    SGZ_TRY(self_fast_image(ctx, cfg, p, g, H, numCh, frames2 != nullptr, 0, ext, dMat.p, &usedTc));
    if (usedTc > 0) {
      std::vector<float> hm((size_t)ext * ext);
      std::vector<int32_t> hi((size_t)ext * ext);
      SGZ_CUDA(cudaMemcpyAsync(hm.data(), dMat.p, hm.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(hi.data(), dImg.p, hi.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
      for (int64_t k = 0; k < nCells; k++) {
        int l = leftIdx[k], r = rightIdx[k];
        if (l > r) {
          if (frames2) continue;          // cross mode: the image holds (left <= right) only
          std::swap(l, r);
        }
        if (sim) sim[k] = hm[(size_t)l * ext + r];
        if (rgb) rgb[k] = hi[(size_t)(ext - 1 - r) * ext + l];
      }
    }
  }
  return SGZ_OK;
}

// openInput, CrossSimilarityImpl.scala:69-82
static void cross_open(int step, int hasStart, int hasStop, int64_t spStart, int64_t spStop, int64_t numFrames,
                       int64_t &start, int64_t &len) {
  const int64_t s = hasStart ? sgz::full_to_feat(spStart, step) : 0;
  const int64_t e = hasStop ? sgz::full_to_feat(spStop, step) : numFrames;
  const int64_t stop = std::min(numFrames, e);
  start = std::max<int64_t>(0, std::min(stop, s));
  len = stop - start;
}

int sgz_cross_num_outputs(const sgz_cross_config *cfg, int64_t nFrames1, int64_t nFrames2, int64_t *nOut) {
  SGZ_REQUIRE(cfg && nOut, "sgz_cross_num_outputs: NULL argument");
  SGZ_REQUIRE(cfg->stepSize > 0, "stepSize must be > 0");
  int64_t st1, l1, st2, l2;
  cross_open(cfg->stepSize, cfg->has1Start, cfg->has1Stop, cfg->span1Start, cfg->span1Stop, nFrames1, st1, l1);
  cross_open(cfg->stepSize, cfg->has2Start, cfg->has2Stop, cfg->span2Start, cfg->span2Stop, nFrames2, st2, l2);
  const int64_t lenA = std::min(l1, l2), lenB = std::max(l1, l2);
  SGZ_REQUIRE(lenA <= sgz::kCrossBuf, "shorter input has %lld feature frames; the reference's 8192-frame buffer "
              "overflows (ArrayIndexOutOfBoundsException)", (long long)lenA);
  SGZ_REQUIRE(lenA > 0 || lenB == 0, "shorter input span is empty (reference: ArithmeticException, % 0)");
  *nOut = lenB > 0 ? 1 + lenB - std::min<int64_t>(lenB, sgz::kCrossBuf) : 0;
  return SGZ_OK;
}

int sgz_cross_run(sgz_ctx *ctx, const sgz_cross_config *cfg, int32_t numCh, const float *norm, const void *frames1,
                  int64_t nFrames1, const void *frames2, int64_t nFrames2, int32_t layout, float *sim, int64_t simCap,
                  int64_t *nOut) {
  using namespace sgz;
  SGZ_REQUIRE(ctx && cfg && frames1 && frames2 && nOut, "sgz_cross_run: NULL argument");
  SGZ_REQUIRE(numCh >= 2 && numCh <= 48, "numCh = numCoeffs + 1 must be in [2,48], got %d", numCh);
  SGZ_REQUIRE(layout >= 0 && layout <= 2, "unknown layout %d", layout);
  SGZ_TRY(ctx->bind());
  int64_t n = 0;
  SGZ_TRY(sgz_cross_num_outputs(cfg, nFrames1, nFrames2, &n));
  *nOut = n;
  if (n == 0) return SGZ_OK;
  SGZ_REQUIRE(!sim || simCap >= n, "sim buffer too small (%lld < %lld)", (long long)simCap, (long long)n);
  int64_t st1, l1, st2, l2;
  cross_open(cfg->stepSize, cfg->has1Start, cfg->has1Stop, cfg->span1Start, cfg->span1Stop, nFrames1, st1, l1);
  cross_open(cfg->stepSize, cfg->has2Start, cfg->has2Stop, cfg->span2Start, cfg->span2Stop, nFrames2, st2, l2);
  // shorter span -> template (afIn1, read completely), longer -> slid over (afIn2); ties go to file 2 (:93-95)
  const bool firstIsTemplate = l1 < l2;
  const void *fA = firstIsTemplate ? frames1 : frames2, *fB = firstIsTemplate ? frames2 : frames1;
  const int64_t nA = firstIsTemplate ? nFrames1 : nFrames2, nB = firstIsTemplate ? nFrames2 : nFrames1;
  const int64_t posA = firstIsTemplate ? st1 : st2, posB = firstIsTemplate ? st2 : st1;
  const int L = (int)(firstIsTemplate ? l1 : l2);
  const int64_t lenB = firstIsTemplate ? l2 : l1;

  // matrixIn (:100-116): normalise, MathUtil.stat per group, ln of the loudness average -- host side, Double
  std::vector<float> planar;
  to_planar(fA, nA, numCh, layout, planar);
  std::vector<float> a((size_t)numCh * L);
  for (int c = 0; c < numCh; c++) {
    const float mn = norm ? norm[2 * c] : 0.f, d = norm ? norm[2 * c + 1] - mn : 1.f;
    for (int i = 0; i < L; i++) {
      const float f = planar[(size_t)c * nA + posA + i];
      a[(size_t)c * L + i] = norm ? (f - mn) / d : f;
    }
  }
  auto stat = [&](int c0, int c1, double &mean, double &sd) {
    double sum = 0.0;
    for (int c = c0; c < c1; c++) for (int i = 0; i < L; i++) sum += a[(size_t)c * L + i];
    const int matSize = L * (c1 - c0);
    mean = sum / matSize;
    sum = 0.0;
    for (int c = c0; c < c1; c++) for (int i = 0; i < L; i++) { double dd = a[(size_t)c * L + i] - mean; sum += dd * dd; }
    sd = sqrt(sum / matSize);
  };
  CrossParams p{};
  stat(0, 1, p.meanT, p.stdT);
  stat(1, numCh, p.meanS, p.stdS);
  {
    double sum = 0.0;
    for (int i = 0; i < L; i++) sum += a[i];
    p.lnAvgIn = log((double)(float)(sum / L));
  }
  // first factor of MathUtil.correlate's products (MathUtil.scala:185-191): a(ch)(i) + aAdd, aAdd = -mean of the group
  std::vector<double> ac(a.size());
  for (int c = 0; c < numCh; c++) {
    const double aAdd = c == 0 ? -p.meanT : -p.meanS;
    for (int i = 0; i < L; i++) ac[(size_t)c * L + i] = (double)a[(size_t)c * L + i] + aAdd;
  }
  DevBuf<float> x, dSim;
  DevBuf<double> dA;
  int64_t stride = 0;
  SGZ_TRY(upload_features(ctx, numCh, norm, fB, nB, layout, posB, lenB, x, stride));
  SGZ_TRY(dA.alloc(a.size()));
  SGZ_TRY(dSim.alloc((size_t)n));
  SGZ_CUDA(cudaMemcpyAsync(dA.p, ac.data(), ac.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  p.x = x.p; p.stride = stride; p.a = dA.p; p.numCh = numCh; p.L = L;
  p.c0 = (int)std::min<int64_t>(lenB, kCrossBuf);
  p.nOut = n;
  p.weight = cfg->temporalWeight; p.maxBoost = cfg->maxBoost;
  p.sim = dSim.p;
  SGZ_TRY(ctx->begin_call());
  k_cross<<<(unsigned)ceil_div<int64_t>(n, 64), 64, 0, ctx->stream>>>(p);
  SGZ_LAUNCH_CHECK(ctx);
  SGZ_TRY(ctx->end_call());
  if (sim) SGZ_CUDA(cudaMemcpyAsync(sim, dSim.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return SGZ_OK;
}

}  // extern "C"
