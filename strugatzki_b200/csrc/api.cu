// api.cu -- the extern "C" boundary (include/strugatzki_b200.h).  Unity build: all kernels are
// included here so that one nvcc invocation produces libsgz_b200.so.
#include "common.cuh"
#include "corr.cuh"
#include "corr_kernel.cuh"
#include "db.cuh"
#include "peaks.cuh"
#include "punchout.cuh"
#include "segm.cuh"
#include "select.cuh"
#include "selfsim.cuh"
#include "stats.cuh"

using namespace sgz;

extern "C" {

int sgz_abi_version(void) { return SGZ_ABI_VERSION; }

const char *sgz_last_error(void) { return err_slot().c_str(); }

int sgz_device_count(int32_t *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount -> %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (count) *count = 0;
    return SGZ_ERR_CUDA;
  }
  int usable = 0;
  for (int d = 0; d < n; d++) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) usable++;
  }
  if (count) *count = usable;
  return SGZ_OK;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
int sgz_ctx_create(int32_t device, sgz_ctx **out) {
  SGZ_REQUIRE(out != nullptr, "sgz_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); strugatzki_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return SGZ_ERR_CUDA;
  }
  SGZ_REQUIRE(device >= 0 && device < n, "sgz_ctx_create: device %d out of range [0,%d)", device, n);
  cudaDeviceProp pr;
  SGZ_CUDA(cudaGetDeviceProperties(&pr, device));
  if (pr.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, pr.major, pr.minor);
    return SGZ_ERR_CUDA;
  }
  sgz_ctx *c = new sgz_ctx();
  c->device = device;
  c->smCount = pr.multiProcessorCount;
  c->smemOptin = pr.sharedMemPerBlockOptin;
  SGZ_CUDA(cudaSetDevice(device));
  SGZ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  SGZ_CUDA(cudaStreamCreateWithFlags(&c->scanStream, cudaStreamNonBlocking));
  SGZ_CUDA(cudaEventCreate(&c->ev0));
  SGZ_CUDA(cudaEventCreate(&c->ev1));
  SGZ_CUDA(cudaEventCreate(&c->evMid));
  *out = c;
  return SGZ_OK;
}

static void ctx_free(sgz_ctx *ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->scanStream);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->evMid);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->scanStream);
  delete ctx;
}

int sgz_ctx_destroy(sgz_ctx *ctx) {
  if (!ctx) return SGZ_OK;
  if (ctx->refs.load() > 0) {   // databases still alive: released by the last sgz_db_destroy
    ctx->zombie = true;
    return SGZ_OK;
  }
  ctx_free(ctx);
  return SGZ_OK;
}

int sgz_ctx_synchronize(sgz_ctx *ctx) {
  SGZ_REQUIRE(ctx, "ctx is NULL");
  SGZ_TRY(ctx->bind());
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return SGZ_OK;
}

void *sgz_ctx_stream(sgz_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int sgz_ctx_last_timing(sgz_ctx *ctx, float *ms, int64_t *launches) {
  SGZ_REQUIRE(ctx, "ctx is NULL");
  if (ms) *ms = ctx->lastMs;
  if (launches) *launches = ctx->lastLaunches;
  return SGZ_OK;
}

int64_t sgz_ctx_launch_count(sgz_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t sgz_self_last_kernel(sgz_ctx *ctx) { return ctx ? ctx->lastSelfKernel : 0; }

int sgz_ctx_trim(sgz_ctx *ctx, int64_t *freedBytes) {
  SGZ_REQUIRE(ctx, "ctx is NULL");
  SGZ_TRY(ctx->bind());
  SGZ_CUDA(cudaDeviceSynchronize());
  size_t n = DevicePool::of(ctx->device).trim();
  if (freedBytes) *freedBytes = (int64_t)n;
  return SGZ_OK;
}

int sgz_measure_peak(sgz_ctx *ctx, int32_t which, double *value) {
  SGZ_REQUIRE(ctx && value, "sgz_measure_peak: NULL argument");
  return measure_peak(ctx, which, value);
}

// ---------------------------------------------------------------------------------------------
// database
// ---------------------------------------------------------------------------------------------
int sgz_db_create(sgz_ctx *ctx, int32_t numCh, const float *norm, sgz_db **out) {
  SGZ_REQUIRE(ctx && out, "sgz_db_create: NULL argument");
  SGZ_REQUIRE(numCh >= 2 && numCh <= 48, "numCh = numCoeffs + 1 must be in [2,48], got %d", numCh);
  SGZ_TRY(ctx->bind());
  sgz_db *db = new sgz_db();
  db->ctx = ctx;
  db->numCh = numCh;
  db->numPairs = (numCh + 1) / 2;
  db->hasNorm = norm != nullptr;
  db->norm.resize((size_t)numCh * 2);
  for (int c = 0; c < numCh; c++) {
    db->norm[2 * c] = norm ? norm[2 * c] : 0.f;
    db->norm[2 * c + 1] = norm ? norm[2 * c + 1] : 1.f;   // (f - 0) / (1 - 0) == f exactly
  }
  int rc = db->dNorm.alloc((size_t)numCh * 2);
  if (rc < 0) { delete db; return rc; }
  SGZ_CUDA(cudaMemcpyAsync(db->dNorm.p, db->norm.data(), db->norm.size() * sizeof(float), cudaMemcpyHostToDevice,
                           ctx->stream));
  SGZ_CUDA(cudaStreamCreateWithFlags(&db->copyStream, cudaStreamNonBlocking));
  for (int i = 0; i < sgz_db::kStageSlots; i++) {
    SGZ_CUDA(cudaEventCreateWithFlags(&db->stageFull[i], cudaEventDisableTiming));
    SGZ_CUDA(cudaEventCreateWithFlags(&db->stageFree[i], cudaEventDisableTiming));
  }
  db->fileStart.push_back(0);
  ctx->refs++;
  *out = db;
  return SGZ_OK;
}

static void db_free(sgz_db *db) {
  sgz_ctx *ctx = db->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(db->copyStream);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->scanStream);
  for (int i = 0; i < sgz_db::kStageSlots; i++) {
    cudaEventDestroy(db->stageFull[i]);
    cudaEventDestroy(db->stageFree[i]);
  }
  for (auto &c : db->chunks) cudaEventDestroy(c.ev);
  cudaStreamDestroy(db->copyStream);
  delete db;
  if (--ctx->refs == 0 && ctx->zombie) ctx_free(ctx);
}

int sgz_db_destroy(sgz_db *db) {
  if (!db) return SGZ_OK;
  if (db->refs.load() > 0) {   // jobs still alive: released by the last sgz_corr_destroy
    db->zombie = true;
    return SGZ_OK;
  }
  db_free(db);
  return SGZ_OK;
}

int sgz_db_reserve(sgz_db *db, int64_t totalFrames, int32_t numFiles) {
  SGZ_REQUIRE(db, "db is NULL");
  SGZ_REQUIRE(!db->finalized, "sgz_db_reserve after finalize");
  SGZ_TRY(db->ctx->bind());
  if (numFiles > 0) db->fileStart.reserve((size_t)numFiles + 1);
  return db_grow(db, totalFrames);
}

static int db_begin_file(sgz_db *db, int64_t nFrames, int64_t *dstFrame) {
  SGZ_REQUIRE(db, "db is NULL");
  if (db->finalized) {
    set_error("database already finalized");
    return SGZ_ERR_STATE;
  }
  SGZ_REQUIRE(nFrames >= 0, "negative frame count");
  SGZ_TRY(db->ctx->bind());
  SGZ_TRY(db_grow(db, db->usedFrames + nFrames));
  *dstFrame = db->usedFrames;
  return SGZ_OK;
}

static int db_mark_chunk(sgz_db *db, int64_t uptoFrame) {
  sgz_db::Chunk c{uptoFrame, nullptr};
  SGZ_CUDA(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
  SGZ_CUDA(cudaEventRecord(c.ev, db->ctx->stream));
  db->chunks.push_back(c);
  db->chunkMark = db->usedFrames;
  return SGZ_OK;
}

static int db_commit_file(sgz_db *db, int64_t nFrames, bool mark = true) {
  db->usedFrames += nFrames;
  db->fileStart.push_back(db->usedFrames);
  // every writer of the pair rows runs on ctx->stream, so an event recorded now covers all frames so far
  if (mark && db->usedFrames - db->chunkMark >= chunk_frames()) SGZ_TRY(db_mark_chunk(db, db->usedFrames));
  return db->numFiles() - 1;
}

// one host -> device copy + the prepare kernels of the files in it
static int db_upload(sgz_db *db, const void *host, size_t bytes, const sgz_db::Pend *files, int nFiles, bool waitCopy) {
  const int s = db->stageIdx;
  db->stageIdx = (s + 1) % sgz_db::kStageSlots;
  if (db->dStage[s].n < bytes) {
    // staging buffer may still be read by an earlier prepare kernel
    if (db->stageUsed[s]) SGZ_CUDA(cudaEventSynchronize(db->stageFree[s]));
    SGZ_TRY(db->dStage[s].alloc(bytes + bytes / 4));
    db->stageUsed[s] = false;
  }
  if (db->stageUsed[s]) SGZ_CUDA(cudaStreamWaitEvent(db->copyStream, db->stageFree[s], 0));
  SGZ_CUDA(cudaMemcpyAsync(db->dStage[s].p, host, bytes, cudaMemcpyHostToDevice, db->copyStream));
  SGZ_CUDA(cudaEventRecord(db->stageFull[s], db->copyStream));
  SGZ_CUDA(cudaStreamWaitEvent(db->ctx->stream, db->stageFull[s], 0));
  for (int i = 0; i < nFiles; i++)
    SGZ_TRY(db_launch_prepare(db, (const float *)(db->dStage[s].p + files[i].off), files[i].layout, files[i].nFrames, files[i].dst));
  SGZ_CUDA(cudaEventRecord(db->stageFree[s], db->ctx->stream));
  db->stageUsed[s] = true;
  if (waitCopy) SGZ_CUDA(cudaEventSynchronize(db->stageFull[s]));
  return SGZ_OK;
}

// uploads the pending run of contiguous HOST_STABLE files (see sgz_db::pend)
static int db_flush(sgz_db *db) {
  if (db->pend.empty()) return SGZ_OK;
  SGZ_TRY(db->ctx->bind());
  SGZ_TRY(db_upload(db, db->pendPtr, db->pendBytes, db->pend.data(), (int)db->pend.size(), false));
  const sgz_db::Pend &last = db->pend.back();
  const int64_t upto = last.dst + last.nFrames;
  db->pend.clear();
  db->pendPtr = nullptr;
  db->pendBytes = 0;
  if (upto - db->chunkMark >= chunk_frames()) {
    SGZ_TRY(db_mark_chunk(db, upto));
    db->chunkMark = upto;
  }
  return SGZ_OK;
}

int sgz_db_add_file(sgz_db *db, const void *frames, int64_t nFrames, int32_t layout) {
  int64_t dst = 0;
  SGZ_TRY(db_begin_file(db, nFrames, &dst));
  SGZ_REQUIRE(frames || nFrames == 0, "frames is NULL");
  const bool hostStable = (layout & SGZ_LAYOUT_HOST_STABLE) != 0;
  layout &= 0xff;
  SGZ_REQUIRE(layout >= 0 && layout <= 2, "unknown layout %d", layout);
  const size_t bytes = (size_t)nFrames * db->numCh * sizeof(float);
  if (hostStable && nFrames > 0) {
    // deferred: joins the run if it continues it in host memory
    const unsigned char *h = (const unsigned char *)frames;
    if (!db->pend.empty() && !(h == db->pendPtr + db->pendBytes && db->pendBytes + bytes <= sgz_db::kRunBytes)) SGZ_TRY(db_flush(db));
    if (db->pend.empty()) db->pendPtr = h;
    db->pend.push_back(sgz_db::Pend{db->pendBytes, nFrames, layout, dst});
    db->pendBytes += bytes;
    return db_commit_file(db, nFrames, false);
  }
  SGZ_TRY(db_flush(db));
  if (nFrames > 0) {
    const sgz_db::Pend one{0, nFrames, layout, dst};
    // the host buffer may be freed on return: pageable copies are staged by the driver before
    // cudaMemcpyAsync returns, pinned copies need the explicit wait
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, frames) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    SGZ_TRY(db_upload(db, frames, bytes, &one, 1, pinned));
  }
  return db_commit_file(db, nFrames);
}

int sgz_db_add_file_device(sgz_db *db, const void *dFrames, int64_t nFrames) {
  int64_t dst = 0;
  SGZ_TRY(db_begin_file(db, nFrames, &dst));
  SGZ_TRY(db_flush(db));
  SGZ_REQUIRE(dFrames || nFrames == 0, "dFrames is NULL");
  SGZ_TRY(db_launch_prepare(db, (const float *)dFrames, SGZ_LAYOUT_INTERLEAVED_LE, nFrames, dst));
  return db_commit_file(db, nFrames);
}

int sgz_db_add_synth(sgz_db *db, uint64_t seed, uint32_t stream, int64_t nFrames, const float *mu,
                     const float *sigma, float floor0) {
  int64_t dst = 0;
  SGZ_TRY(db_begin_file(db, nFrames, &dst));
  SGZ_TRY(db_flush(db));
  SGZ_REQUIRE(mu && sigma, "mu / sigma is NULL");
  if (nFrames > 0) {
    DevBuf<float> ms;
    SGZ_TRY(ms.alloc((size_t)db->numCh * 2));
    SGZ_CUDA(cudaMemcpyAsync(ms.p, mu, db->numCh * sizeof(float), cudaMemcpyHostToDevice, db->ctx->stream));
    SGZ_CUDA(cudaMemcpyAsync(ms.p + db->numCh, sigma, db->numCh * sizeof(float), cudaMemcpyHostToDevice,
                             db->ctx->stream));
    dim3 grid((unsigned)std::min<int64_t>(ceil_div<int64_t>(nFrames, 256), 4096), (unsigned)db->numPairs);
    k_db_synth<<<grid, 256, 0, db->ctx->stream>>>(db->dData.p, db->capFrames, dst, nFrames, db->numCh, seed, stream,
                                                  ms.p, ms.p + db->numCh, floor0, db->dNorm.p);
    SGZ_LAUNCH_CHECK(db->ctx);
    SGZ_CUDA(cudaStreamSynchronize(db->ctx->stream));  // ms goes out of scope
  }
  return db_commit_file(db, nFrames);
}

int sgz_db_add_synth_many(sgz_db *db, uint64_t seed, uint32_t firstStream, int32_t numFiles, int64_t nFramesEach,
                          const float *mu, const float *sigma, float floor0) {
  SGZ_REQUIRE(numFiles >= 0 && nFramesEach >= 0, "negative file / frame count");
  int64_t dst = 0;
  SGZ_TRY(db_begin_file(db, (int64_t)numFiles * nFramesEach, &dst));
  SGZ_TRY(db_flush(db));
  SGZ_REQUIRE(mu && sigma, "mu / sigma is NULL");
  const int first = db->numFiles();
  if (numFiles > 0 && nFramesEach > 0) {
    DevBuf<float> ms;
    SGZ_TRY(ms.alloc((size_t)db->numCh * 2));
    SGZ_CUDA(cudaMemcpyAsync(ms.p, mu, db->numCh * sizeof(float), cudaMemcpyHostToDevice, db->ctx->stream));
    SGZ_CUDA(cudaMemcpyAsync(ms.p + db->numCh, sigma, db->numCh * sizeof(float), cudaMemcpyHostToDevice,
                             db->ctx->stream));
    const int64_t total = (int64_t)numFiles * nFramesEach;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)db->ctx->smCount * 32),
              (unsigned)db->numPairs);
    k_db_synth_many<<<grid, 256, 0, db->ctx->stream>>>(db->dData.p, db->capFrames, dst, nFramesEach, numFiles, db->numCh,
                                                       seed, firstStream, ms.p, ms.p + db->numCh, floor0, db->dNorm.p);
    SGZ_LAUNCH_CHECK(db->ctx);
    SGZ_CUDA(cudaStreamSynchronize(db->ctx->stream));  // ms goes out of scope
  }
  for (int f = 0; f < numFiles; f++) SGZ_TRY(db_commit_file(db, nFramesEach));
  return first;
}

int sgz_db_patch(sgz_db *db, int32_t file, int64_t frameOff, const float *frames, int64_t n) {
  SGZ_REQUIRE(db && frames, "NULL argument");
  SGZ_REQUIRE(file >= 0 && file < db->numFiles(), "file index %d out of range", file);
  int64_t len = db->fileStart[file + 1] - db->fileStart[file];
  SGZ_REQUIRE(frameOff >= 0 && n >= 0 && frameOff + n <= len, "patch [%lld,+%lld) outside file of %lld frames",
              (long long)frameOff, (long long)n, (long long)len);
  SGZ_TRY(db->ctx->bind());
  if (n == 0) return SGZ_OK;
  SGZ_TRY(db_flush(db));
  DevBuf<float> tmp;
  SGZ_TRY(tmp.alloc((size_t)n * db->numCh));
  SGZ_CUDA(cudaMemcpyAsync(tmp.p, frames, (size_t)n * db->numCh * sizeof(float), cudaMemcpyHostToDevice,
                           db->ctx->stream));
  SGZ_TRY(db_launch_prepare(db, tmp.p, SGZ_LAYOUT_INTERLEAVED_LE, n, db->fileStart[file] + frameOff));
  SGZ_CUDA(cudaStreamSynchronize(db->ctx->stream));
  db->planesUpto = 0;   // the FP16 planes of the tensor-core K1 are rebuilt by the next search
  return SGZ_OK;
}

// everything of finalize that is stream ordered: no host wait for the uploads
static int db_finalize_enqueue(sgz_db *db) {
  SGZ_TRY(db_flush(db));
  SGZ_TRY(db_grow(db, db->usedFrames));  // guarantees the slack even for an empty DB
  SGZ_CUDA(cudaMemset2DAsync(db->dData.p + db->usedFrames, (size_t)db->capFrames * sizeof(float2), 0,
                             (size_t)kDbSlack * sizeof(float2), (size_t)db->numPairs, db->ctx->stream));
  SGZ_TRY(db->dFileStart.alloc(db->fileStart.size()));
  // the file table goes through the (idle) scan stream so that it does not queue behind uploads in flight
  SGZ_CUDA(cudaMemcpyAsync(db->dFileStart.p, db->fileStart.data(), db->fileStart.size() * sizeof(int64_t),
                           cudaMemcpyHostToDevice, db->ctx->scanStream));
  SGZ_CUDA(cudaStreamSynchronize(db->ctx->scanStream));
  SGZ_TRY(db_mark_chunk(db, INT64_MAX));   // "everything, including the slack"
  db->finalized = true;
  return SGZ_OK;
}

// host wait until every upload has landed; drops the progress markers and the staging ring
static int db_wait_resident(sgz_db *db) {
  SGZ_CUDA(cudaStreamSynchronize(db->copyStream));
  SGZ_CUDA(cudaStreamSynchronize(db->ctx->stream));
  for (auto &c : db->chunks) cudaEventDestroy(c.ev);
  db->chunks.clear();
  for (int i = 0; i < sgz_db::kStageSlots; i++) {
    db->dStage[i].release();
    db->stageUsed[i] = false;
  }
  return SGZ_OK;
}

int sgz_db_finalize(sgz_db *db) {
  SGZ_REQUIRE(db, "db is NULL");
  SGZ_TRY(db->ctx->bind());
  if (!db->finalized) SGZ_TRY(db_finalize_enqueue(db));
  return db_wait_resident(db);   // also the explicit wait after sgz_db_finalize_async
}

int sgz_db_finalize_async(sgz_db *db) {
  SGZ_REQUIRE(db, "db is NULL");
  SGZ_TRY(db->ctx->bind());
  if (!db->finalized) SGZ_TRY(db_finalize_enqueue(db));
  return SGZ_OK;
}

int sgz_db_stats(sgz_db *db, double *out, double *perFileOut) {
  SGZ_REQUIRE(db && out, "sgz_db_stats: NULL argument");
  if (!db->finalized) { set_error("sgz_db_stats: database not finalized"); return SGZ_ERR_STATE; }
  SGZ_REQUIRE(!db->hasNorm, "sgz_db_stats works on RAW features: create the database with norm = NULL");
  sgz_ctx *ctx = db->ctx;
  SGZ_TRY(ctx->bind());
  SGZ_TRY(db_wait_resident(db));
  const int nf = db->numFiles(), nc = db->numCh;
  SGZ_REQUIRE(nf > 0, "sgz_db_stats: empty database");
  DevBuf<float> dMins, dMaxs;
  DevBuf<double> dSkews, dPer;
  DevBuf<int32_t> dHist;
  SGZ_TRY(dMins.alloc((size_t)nf * nc));
  SGZ_TRY(dMaxs.alloc((size_t)nf * nc));
  SGZ_TRY(dSkews.alloc((size_t)nf * nc));
  SGZ_TRY(dPer.alloc((size_t)nf * nc * 2));
  const int batch = std::min(nf, 1024);
  SGZ_TRY(dHist.alloc((size_t)batch * nc * kStatBins));
  StatsParams p{};
  p.data = db->dData.p; p.rowStride = db->capFrames; p.fileStart = db->dFileStart.p;
  p.numFiles = nf; p.numCh = nc; p.numPairs = db->numPairs;
  p.mins = dMins.p; p.maxs = dMaxs.p; p.skews = dSkews.p; p.hist = dHist.p; p.perFile = dPer.p;
  SGZ_TRY(ctx->begin_call());
  k_stats_minmax<<<ceil_div(nf * db->numPairs, kMmWarps), 32 * kMmWarps, 0, ctx->stream>>>(p);
  SGZ_LAUNCH_CHECK(ctx);
  for (int f0 = 0; f0 < nf; f0 += batch) {
    p.file0 = f0;
    p.fileCount = std::min(batch, nf - f0);
    int64_t longest = 0;
    for (int f = f0; f < f0 + p.fileCount; f++) longest = std::max(longest, db->fileStart[f + 1] - db->fileStart[f]);
    SGZ_CUDA(cudaMemsetAsync(dHist.p, 0, (size_t)p.fileCount * nc * kStatBins * sizeof(int32_t), ctx->stream));
    if (longest > 0) {
      dim3 grid((unsigned)ceil_div<int64_t>(longest, kStatSeg), (unsigned)db->numPairs, (unsigned)p.fileCount);
      k_stats_hist<<<grid, 256, 0, ctx->stream>>>(p);
      SGZ_LAUNCH_CHECK(ctx);
    }
    k_stats_pctl<<<ceil_div(p.fileCount * nc, 4), 128, 0, ctx->stream>>>(p);   // one warp per (file, channel)
    SGZ_LAUNCH_CHECK(ctx);
  }
  SGZ_TRY(ctx->end_call());
  std::vector<double> per((size_t)nf * nc * 2);
  SGZ_CUDA(cudaMemcpyAsync(per.data(), dPer.p, per.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
  // body(), FeatureStatsImpl.scala:30-53: math.min / math.max over the files, in order (NaN propagates)
  auto jmin = [](double a, double b) { return (a != a || b != b) ? (double)NAN : (a < b ? a : b); };
  auto jmax = [](double a, double b) { return (a != a || b != b) ? (double)NAN : (a > b ? a : b); };
  for (int f = 0; f < nf; f++)
    for (int c = 0; c < nc; c++) {
      const double lo = per[((size_t)f * nc + c) * 2], hi = per[((size_t)f * nc + c) * 2 + 1];
      out[2 * c] = f == 0 ? lo : jmin(out[2 * c], lo);
      out[2 * c + 1] = f == 0 ? hi : jmax(out[2 * c + 1], hi);
    }
  if (perFileOut) memcpy(perFileOut, per.data(), per.size() * sizeof(double));
  return SGZ_OK;
}

int sgz_db_info(sgz_db *db, int32_t *numFiles, int64_t *totalFrames, int32_t *numCh) {
  SGZ_REQUIRE(db, "db is NULL");
  if (numFiles) *numFiles = db->numFiles();
  if (totalFrames) *totalFrames = db->usedFrames;
  if (numCh) *numCh = db->numCh;
  return SGZ_OK;
}

int sgz_db_file_frames(sgz_db *db, int32_t file, int64_t *nFrames) {
  SGZ_REQUIRE(db && nFrames, "NULL argument");
  SGZ_REQUIRE(file >= 0 && file < db->numFiles(), "file index %d out of range", file);
  *nFrames = db->fileStart[file + 1] - db->fileStart[file];
  return SGZ_OK;
}

int sgz_db_read(sgz_db *db, int32_t file, int64_t frameOff, int64_t n, float *out) {
  SGZ_REQUIRE(db && out, "NULL argument");
  SGZ_REQUIRE(file >= 0 && file < db->numFiles(), "file index %d out of range", file);
  int64_t len = db->fileStart[file + 1] - db->fileStart[file];
  SGZ_REQUIRE(frameOff >= 0 && n >= 0 && frameOff + n <= len, "read outside file");
  SGZ_TRY(db->ctx->bind());
  if (n == 0) return SGZ_OK;
  SGZ_TRY(db_flush(db));
  DevBuf<float> tmp;
  SGZ_TRY(tmp.alloc((size_t)n * db->numCh));
  k_db_gather<<<(unsigned)ceil_div<int64_t>(n * db->numCh, 256), 256, 0, db->ctx->stream>>>(
      db->dData.p, db->capFrames, db->fileStart[file] + frameOff, n, db->numCh, tmp.p);
  SGZ_LAUNCH_CHECK(db->ctx);
  SGZ_CUDA(cudaMemcpyAsync(out, tmp.p, (size_t)n * db->numCh * sizeof(float), cudaMemcpyDeviceToHost,
                           db->ctx->stream));
  SGZ_CUDA(cudaStreamSynchronize(db->ctx->stream));
  return SGZ_OK;
}

// ---------------------------------------------------------------------------------------------
// FeatureCorrelation
// ---------------------------------------------------------------------------------------------
static void to_planar(const void *src, int64_t nFrames, int numCh, int layout, std::vector<float> &out) {
  out.resize((size_t)nFrames * numCh);
  const float *s = (const float *)src;
  if (layout == SGZ_LAYOUT_PLANAR_LE) {
    memcpy(out.data(), s, out.size() * sizeof(float));
    return;
  }
  for (int64_t t = 0; t < nFrames; t++)
    for (int c = 0; c < numCh; c++) {
      float v = s[t * numCh + c];
      if (layout == SGZ_LAYOUT_INTERLEAVED_BE) {
        uint32_t b;
        memcpy(&b, &v, 4);
        b = __builtin_bswap32(b);
        memcpy(&v, &b, 4);
      }
      out[(size_t)c * nFrames + t] = v;
    }
}

int sgz_corr_create(sgz_db *db, const sgz_corr_config *cfg, const void *input, int64_t inputFrames, int32_t layout,
                    sgz_corr **out) {
  SGZ_REQUIRE(db && cfg && input && out, "sgz_corr_create: NULL argument");
  *out = nullptr;
  if (!db->finalized) {
    set_error("sgz_corr_create: database not finalized");
    return SGZ_ERR_STATE;
  }
  SGZ_REQUIRE(cfg->stepSize > 0, "stepSize must be > 0");
  SGZ_REQUIRE(layout >= 0 && layout <= 2, "unknown layout %d", layout);
  // numPerFile = 0 is degenerate in the reference (addMatch's collapse branch has no size check, so matches leak
  // into a queue that is supposed to hold none, FeatureCorrelationImpl.scala:136-143); rejected instead of emulated
  SGZ_REQUIRE(cfg->numPerFile >= 1, "numPerFile must be >= 1, got %d", cfg->numPerFile);
  SGZ_TRY(db->ctx->bind());
  sgz_corr *job = new sgz_corr();
  job->db = db;
  job->ctx = db->ctx;
  job->cfg = *cfg;
  job->step = cfg->stepSize;
  job->hasOut = cfg->hasPunchOut != 0;
  job->minPunchF = full_to_feat(cfg->minPunch, job->step);
  job->maxPunchF = full_to_feat(cfg->maxPunch, job->step);
  std::vector<float> planar;
  to_planar(input, inputFrames, db->numCh, layout, planar);
  // uploads still in flight (sgz_db_finalize_async): the taps must not queue behind them on the context stream
  const bool streaming = !db->chunks.empty();
  cudaStream_t tapStream = streaming ? db->ctx->scanStream : db->ctx->stream;
  int rc = prepare_query(db, planar.data(), inputFrames, cfg->punchInStart, cfg->punchInStop, cfg->punchInWeight,
                         job->step, job->qin, tapStream);
  if (rc == SGZ_OK && job->hasOut)
    rc = prepare_query(db, planar.data(), inputFrames, cfg->punchOutStart, cfg->punchOutStop, cfg->punchOutWeight,
                       job->step, job->qout, tapStream);
  if (rc == SGZ_OK && streaming && cudaStreamSynchronize(tapStream) != cudaSuccess) {
    set_error("taps upload failed");
    rc = SGZ_ERR_CUDA;
  }
  if (rc == SGZ_OK) {
    int wq = std::max(job->qin.Wq, job->hasOut ? job->qout.Wq : 0);
    job->ntg = pick_ntg(job->ctx, db->numPairs, wq);
    if (const char *e = getenv("SGZ_CORR_NSLOT")) job->nslot = atoi(e) == 2 ? 2 : 3;   // tuning override
    if (job->ntg > 0 && corr_smem_layout(job->ntg, db->numPairs, wq, job->nslot).total > job->ctx->smemOptin) job->nslot = 2;
    if (job->ntg > 0 && corr_smem_layout(job->ntg, db->numPairs, wq, job->nslot).total > job->ctx->smemOptin) job->ntg = 0;
    // windows that do not fit the FFMA2 kernel's shared-memory tile (> ~1500 frames) run on the tensor-core kernel alone
    // (corr_tc2.cuh cuts long windows into passes: no length limit short of the zero slack behind the database)
    const bool t2ok = job->qin.dT2Taps.p && (!job->hasOut || job->qout.dT2Taps.p) &&
                      !(getenv("SGZ_CORR_TC") && atoi(getenv("SGZ_CORR_TC")) == 0) &&
                      !(getenv("SGZ_CORR_TC2") && atoi(getenv("SGZ_CORR_TC2")) == 0);
    if (job->ntg == 0 && !t2ok) {
      set_error("punch window of %d feature frames does not fit any scan kernel", wq);
      rc = SGZ_ERR_INVALID;
    }
  }
  if (rc != SGZ_OK) { delete job; return rc; }
  job->numTiles = job->ntg > 0 ? ceil_div<int64_t>(std::max<int64_t>(db->usedFrames, 1), (int64_t)kR * job->ntg) : 0;
  job->numTilesTc = ceil_div<int64_t>(std::max<int64_t>(db->usedFrames, 1), (int64_t)kTcTile);
  {
    // K1 on the tensor cores (corr_tc.cuh) wherever it applies (<= 14 channels, window <= 256 frames); the FFMA2
    // kernel covers the rest and streaming scans.  SGZ_CORR_TC=0 forces the FFMA2 kernel.
    const char *e = getenv("SGZ_CORR_TC");
    job->useTc = !(e && atoi(e) == 0) && job->qin.dTcTaps.p && (!job->hasOut || job->qout.dTcTaps.p);
    // the N = 64 tensor-core kernel fed by bulk copies (corr_tc2.cuh) is the default for resident AND streaming scans;
    // SGZ_CORR_TC2=0 falls back to the round-1 kernels above
    const char *e2 = getenv("SGZ_CORR_TC2");
    job->useT2 = !(e && atoi(e) == 0) && !(e2 && atoi(e2) == 0) && job->qin.dT2Taps.p && (!job->hasOut || job->qout.dT2Taps.p);
    job->numTilesT2 = ceil_div<int64_t>(std::max<int64_t>(db->usedFrames, 1), (int64_t)kT2Tile);
  }
  job->numOffsets = valid_offsets(db, job->qin.W, job->hasOut ? job->minPunchF : 0);
  db->refs++;
  *out = job;
  return SGZ_OK;
}

int sgz_corr_destroy(sgz_corr *job) {
  if (!job) return SGZ_OK;
  if (job->worker.joinable()) {
    job->abortFlag = 1;
    job->worker.join();
  }
  sgz_db *db = job->db;
  cudaSetDevice(job->ctx->device);
  cudaStreamSynchronize(job->ctx->stream);
  delete job;
  if (--db->refs == 0 && db->zombie) db_free(db);
  return SGZ_OK;
}

static int enqueue_summary_download(sgz_corr *job, cudaStream_t st);

int sgz_corr_scan(sgz_corr *job) {
  SGZ_REQUIRE(job, "job is NULL");
  sgz_ctx *ctx = job->ctx;
  sgz_db *db = job->db;
  SGZ_TRY(ctx->bind());
  if (job->abortFlag) return SGZ_ERR_ABORTED;
  job->tailMs = 0.f;
  size_t n = (size_t)job->numTiles * kR * job->ntg;
  const bool t2 = job->useT2;
  const bool tc = !t2 && job->useTc && db->chunks.empty();   // round-1 kernels: a streaming scan takes the FFMA2 path
  if (t2) {
    n = std::max(n, (size_t)job->numTilesT2 * kT2Tile);
    if (!job->dTileFileT2.p) {
      // built on the device from the file table (151 KB for the 1000 h database: a host copy of that size would queue
      // behind the uploads of a streaming scan)
      SGZ_TRY(job->dTileFileT2.alloc((size_t)job->numTilesT2 + 1));
      cudaStream_t tst = db->chunks.empty() ? ctx->stream : ctx->scanStream;
      k_t2_tile_files<<<(unsigned)ceil_div<int64_t>(job->numTilesT2 + 1, 256), 256, 0, tst>>>(
          db->dFileStart.p, db->numFiles(), db->usedFrames, job->numTilesT2, job->dTileFileT2.p);
      SGZ_LAUNCH_CHECK(ctx);
    }
    SGZ_TRY(job->dFixCount.alloc(2));
    SGZ_TRY(job->dFixList[0].alloc(kFixCap));
    if (job->hasOut) SGZ_TRY(job->dFixList[1].alloc(kFixCap));
  }
  if (tc) {
    n = std::max(n, (size_t)job->numTilesTc * kTcTile);
    if (!job->dTileFile.p) {
      std::vector<int32_t> tf((size_t)job->numTilesTc + 1);
      int f = 0;
      const int nf = db->numFiles();
      for (int64_t t = 0; t <= job->numTilesTc; t++) {
        const int64_t g = std::min<int64_t>(t * kTcTile, std::max<int64_t>(db->usedFrames - 1, 0));
        while (f + 1 < nf && db->fileStart[f + 1] <= g) f++;
        tf[(size_t)t] = f;
      }
      SGZ_TRY(job->dTileFile.alloc(tf.size()));
      SGZ_CUDA(cudaMemcpy(job->dTileFile.p, tf.data(), tf.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
  }
  SGZ_TRY(job->simIn.alloc(n));
  // the tensor-core scan writes no boost curve (BoostSrc, corr_fix.cuh); punch-out searches get one from k_boost_all
  if (!t2 || job->hasOut) SGZ_TRY(job->boostIn.alloc(n));
  SGZ_TRY(job->dFileMax.alloc((size_t)std::max(db->numFiles(), 1)));
  {
    static const bool directOff = getenv("SGZ_DIRECT_SELECT") && atoi(getenv("SGZ_DIRECT_SELECT")) == 0;   // developer knob
    job->direct = t2 && !job->hasOut && job->cfg.numPerFile == 1 && !directOff;
    job->keysCached = false;
    static const bool refineOff = getenv("SGZ_REFINE") && atoi(getenv("SGZ_REFINE")) == 0;   // developer knob
    job->refine = t2 && !job->hasOut && !refineOff;
    // filter mode (opt-in): one FP16 product instead of three in K1, the exact re-evaluation makes up for it -- only where
    // the result is decided by the re-evaluated offsets alone (numPerFile = 1 punch-in searches on a resident database)
    static const bool fastOn = getenv("SGZ_FAST") && atoi(getenv("SGZ_FAST")) == 1;   // developer knob
    job->fast = fastOn && job->refine && job->cfg.numPerFile == 1 && db->chunks.empty() &&
                t2_geom(job->qin.W, ctx->smemOptin).NP == 1;
    if (job->refine) {
      SGZ_TRY(job->dRefList.alloc(kRefineCap));
      SGZ_TRY(job->dRefCount.alloc(2));
      SGZ_TRY(job->dRefCand.alloc((size_t)std::max(db->numFiles(), 1)));
      SGZ_TRY(job->dRefThr.alloc(1));
      SGZ_TRY(job->dFileMaxExact.alloc((size_t)std::max(db->numFiles(), 1)));
    }
    if (job->direct) {
      SGZ_TRY(job->dFileNaN.alloc((size_t)std::max(db->numFiles(), 1)));
      SGZ_TRY(job->dFileBoost.alloc((size_t)std::max(db->numFiles(), 1)));
    }
  }
  if (job->hasOut) {
    SGZ_TRY(job->simOut.alloc(n));
    SGZ_TRY(job->boostOut.alloc(n));
    SGZ_TRY(job->dFileMaxOut.alloc((size_t)std::max(db->numFiles(), 1)));
  }
  const int tailIn = job->hasOut ? job->minPunchF : 0;
  if (!db->chunks.empty()) {
    // ---- streaming scan: the database is still being uploaded (sgz_db_finalize_async).  K1 runs on its own
    // stream, range by range behind the upload markers, and leaves a few SMs to the prepare kernels. ----
    cudaStream_t ss = ctx->scanStream;
    job->summaryPrefetched = false;
    const int64_t launches0 = ctx->launches;
    const int64_t T = (int64_t)kR * job->ntg;
    const int wqMax = std::max(job->qin.Wq, job->hasOut ? job->qout.Wq : 0);
    const int spare = ctx->smCount > 32 ? 8 : 0;
    SGZ_CUDA(cudaEventRecord(ctx->ev0, ss));
    SGZ_CUDA(cudaMemsetAsync(job->dFileMax.p, 0, job->dFileMax.n * sizeof(unsigned long long), ss));
    if (job->hasOut)
      SGZ_CUDA(cudaMemsetAsync(job->dFileMaxOut.p, 0, job->dFileMaxOut.n * sizeof(unsigned long long), ss));
    if (t2) SGZ_CUDA(cudaMemsetAsync(job->dFixCount.p, 0, 2 * sizeof(uint32_t), ss));
    if (job->direct) SGZ_CUDA(cudaMemsetAsync(job->dFileNaN.p, 0, job->dFileNaN.n * sizeof(uint32_t), ss));
    int64_t done = 0;
    for (const auto &c : db->chunks) {
      if (db->usedFrames == 0) continue;
      if (t2) {
        // planes of the whole 2048-frame blocks that have landed, then the tiles whose operands (8192 frames + 4 rows) and
        // window statistics (+ W + 16 frames) lie inside them
        const bool last = c.uptoFrame == INT64_MAX;
        const int64_t ready = last ? INT64_MAX : c.uptoFrame / kPlaneFrames * kPlaneFrames;
        const int64_t end = last ? job->numTilesT2
                                 : std::min<int64_t>(job->numTilesT2, std::max<int64_t>(ready - std::max(320, wqMax + 384), 0) / kT2Tile);
        if (end <= done && !last) continue;
        SGZ_CUDA(cudaStreamWaitEvent(ss, c.ev, 0));
        SGZ_TRY(db_ensure_planes(db, last ? -1 : c.uptoFrame, ss));
        SGZ_TRY(run_scan_t2(job, job->qin, 0, tailIn, job->simIn.p, job->dFileMax.p, done, end, ss, spare));
        if (job->hasOut)
          SGZ_TRY(run_scan_t2(job, job->qout, 1, 0, job->simOut.p, job->dFileMaxOut.p, done, end, ss, spare));
        done = std::max(done, end);
        continue;
      }
      // a tile reads frames [tile*T, (tile+1)*T + Wq)
      const int64_t end = c.uptoFrame == INT64_MAX ? job->numTiles
                                                   : std::min<int64_t>(job->numTiles, std::max<int64_t>(c.uptoFrame - wqMax, 0) / T);
      if (end <= done) continue;
      SGZ_CUDA(cudaStreamWaitEvent(ss, c.ev, 0));
      SGZ_TRY(run_scan_one(job, job->qin, tailIn, job->simIn.p, job->boostIn.p, job->dFileMax.p, done, end, ss, spare));
      if (job->hasOut)
        SGZ_TRY(run_scan_one(job, job->qout, 0, job->simOut.p, job->boostOut.p, job->dFileMaxOut.p, done, end, ss, spare));
      done = end;
    }
    SGZ_CUDA(cudaStreamWaitEvent(ss, db->chunks.back().ev, 0));   // file table for the row maxima / later kernels
    if (t2 && db->usedFrames > 0) {
      SGZ_TRY(run_fixup(job, job->qin, 0, tailIn, job->simIn.p, job->dFileMax.p, ss));
      if (job->refine) SGZ_TRY(run_refine(job, ss));
      if (job->hasOut) SGZ_TRY(run_fixup(job, job->qout, 1, 0, job->simOut.p, job->dFileMaxOut.p, ss));
      if (job->hasOut) SGZ_TRY(run_boost_curves(job, ss));
      if (job->direct) {
        SGZ_TRY(run_filemax_boost(job, ss));
      }
    }
    if (job->hasOut && db->usedFrames > 0) {
      SGZ_TRY(job->rowMaxOut.alloc((size_t)db->usedFrames));
      SGZ_CUDA(launch_row_max_out(ss, job->simOut.p, db->dFileStart.p, db->numFiles(), db->usedFrames, job->qin.W,
                                  job->qout.W, job->minPunchF, job->maxPunchF, job->rowMaxOut.p));
      ctx->launches++;
    }
    SGZ_CUDA(cudaEventRecord(ctx->ev1, ss));
    SGZ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev1, 0));
    SGZ_CUDA(cudaEventSynchronize(ctx->ev1));
    SGZ_CUDA(cudaEventElapsedTime(&ctx->lastMs, ctx->ev0, ctx->ev1));
    ctx->lastLaunches = ctx->launches - launches0;
    SGZ_TRY(db_wait_resident(db));
  } else {
    // once per database (a property of the database like its normalisation, not of the search): FP16 planes + frame sums
    if (t2 && db->usedFrames > 0) SGZ_TRY(db_ensure_planes(db, -1, ctx->stream));
    bool midRecorded = false;
    SGZ_TRY(ctx->begin_call());
    SGZ_CUDA(cudaMemsetAsync(job->dFileMax.p, 0, job->dFileMax.n * sizeof(unsigned long long), ctx->stream));
    if (job->hasOut)
      SGZ_CUDA(cudaMemsetAsync(job->dFileMaxOut.p, 0, job->dFileMaxOut.n * sizeof(unsigned long long), ctx->stream));
    if (t2) SGZ_CUDA(cudaMemsetAsync(job->dFixCount.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    if (job->direct) SGZ_CUDA(cudaMemsetAsync(job->dFileNaN.p, 0, job->dFileNaN.n * sizeof(uint32_t), ctx->stream));
    if (db->usedFrames > 0) {
      if (t2) {
        SGZ_TRY(run_scan_t2(job, job->qin, 0, tailIn, job->simIn.p, job->dFileMax.p, 0, job->numTilesT2,
                            ctx->stream, 0));
        if (!job->hasOut) { SGZ_CUDA(cudaEventRecord(ctx->evMid, ctx->stream)); midRecorded = true; }
        SGZ_TRY(run_fixup(job, job->qin, 0, tailIn, job->simIn.p, job->dFileMax.p, ctx->stream));
        if (job->refine) SGZ_TRY(run_refine(job, ctx->stream));
        if (job->direct) {
          SGZ_TRY(run_filemax_boost(job, ctx->stream));
        }
      } else if (tc) SGZ_TRY(run_scan_tc(job, job->qin, tailIn, job->simIn.p, job->boostIn.p, job->dFileMax.p, ctx->stream));
      else
        SGZ_TRY(run_scan_one(job, job->qin, tailIn, job->simIn.p, job->boostIn.p, job->dFileMax.p, 0, job->numTiles,
                             ctx->stream, 0));
      if (job->hasOut) {
        if (t2) {
          SGZ_TRY(run_scan_t2(job, job->qout, 1, 0, job->simOut.p, job->dFileMaxOut.p, 0, job->numTilesT2,
                              ctx->stream, 0));
          SGZ_TRY(run_fixup(job, job->qout, 1, 0, job->simOut.p, job->dFileMaxOut.p, ctx->stream));
          SGZ_TRY(run_boost_curves(job, ctx->stream));
        } else if (tc) SGZ_TRY(run_scan_tc(job, job->qout, 0, job->simOut.p, job->boostOut.p, job->dFileMaxOut.p, ctx->stream));
        else
          SGZ_TRY(run_scan_one(job, job->qout, 0, job->simOut.p, job->boostOut.p, job->dFileMaxOut.p, 0, job->numTiles,
                               ctx->stream, 0));
        SGZ_TRY(job->rowMaxOut.alloc((size_t)db->usedFrames));
        SGZ_CUDA(launch_row_max_out(ctx->stream, job->simOut.p, db->dFileStart.p, db->numFiles(), db->usedFrames,
                                    job->qin.W, job->qout.W, job->minPunchF, job->maxPunchF, job->rowMaxOut.p));
        ctx->launches++;
      }
    }
    job->summaryPrefetched = false;
    SGZ_TRY(ctx->end_call_async());
    SGZ_TRY(enqueue_summary_download(job, ctx->stream));     // the per-file results ride on the scan's own wait
    SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
    SGZ_TRY(ctx->collect_call());
    job->summaryPrefetched = true;
    if (midRecorded) {   // punch-in scan on the tensor cores: the K1 launch on its own, what follows it counts as selection
      float k1 = 0.f;
      SGZ_CUDA(cudaEventElapsedTime(&k1, ctx->ev0, ctx->evMid));
      job->tailMs = ctx->lastMs - k1;
      ctx->lastMs = k1;
    }
  }
  job->scanMs = ctx->lastMs;
  job->scanLaunches = ctx->lastLaunches;
  job->scanned = true;
  job->progress = 0.8f;
  // reset selection state
  job->allPrio.clear();
  job->nextFile = 0;
  job->finished = false;
  job->globalSet = false;
  job->selectMs = job->tailMs;
  return SGZ_OK;
}

// downloads of the per-file results of a scan into the job's pinned scratch, enqueued on `st` (no wait):
// [keys in][keys out][NaN flags][boosts], nf entries each
static int enqueue_summary_download(sgz_corr *job, cudaStream_t st) {
  const int nf = job->db->numFiles();
  SGZ_TRY(job->pin((size_t)std::max(nf, 1) * (2 * sizeof(unsigned long long) + 8)));
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(job->hPin), *keysOut = keys + std::max(nf, 1);
  uint32_t *nanFlags = reinterpret_cast<uint32_t *>(keysOut + std::max(nf, 1));
  float *boosts = reinterpret_cast<float *>(nanFlags + std::max(nf, 1));
  if (nf == 0) return SGZ_OK;
  if (job->direct && job->db->usedFrames > 0) {
    SGZ_CUDA(cudaMemcpyAsync(nanFlags, job->dFileNaN.p, nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SGZ_CUDA(cudaMemcpyAsync(boosts, job->dFileBoost.p, nf * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  SGZ_CUDA(cudaMemcpyAsync(keys, job->dFileMax.p, nf * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (job->hasOut)
    SGZ_CUDA(cudaMemcpyAsync(keysOut, job->dFileMaxOut.p, nf * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  return SGZ_OK;
}

int sgz_corr_local_summary(sgz_corr *job, sgz_file_summary *out, int32_t cap, int32_t *n) {
  SGZ_REQUIRE(job && n, "NULL argument");
  if (!job->scanned) { set_error("sgz_corr_local_summary before scan"); return SGZ_ERR_STATE; }
  sgz_db *db = job->db;
  SGZ_TRY(job->ctx->bind());
  const int nf = db->numFiles();
  *n = nf;
  if (!out) return SGZ_OK;
  SGZ_REQUIRE(cap >= nf, "summary buffer too small (%d < %d)", cap, nf);
  if (!job->summaryPrefetched) {     // (a resident scan has fetched them behind its kernels: one wait less per search)
    SGZ_TRY(enqueue_summary_download(job, job->ctx->stream));
    if (nf > 0) SGZ_CUDA(cudaStreamSynchronize(job->ctx->stream));
    job->summaryPrefetched = true;     // the pinned copy stays valid until the next selection round uses the scratch
  }
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(job->hPin), *keysOut = keys + std::max(nf, 1);
  uint32_t *nanFlags = reinterpret_cast<uint32_t *>(keysOut + std::max(nf, 1));
  float *boosts = reinterpret_cast<float *>(nanFlags + std::max(nf, 1));
  const bool direct = job->direct && db->usedFrames > 0;
  if (!job->hasOut) memset(keysOut, 0, (size_t)std::max(nf, 1) * sizeof(unsigned long long));
  if (direct) {
    job->hKeys.assign(keys, keys + nf);
    job->hNaN.assign(nanFlags, nanFlags + nf);
    job->hBoost.assign(boosts, boosts + nf);
    job->keysCached = true;
  }
  const int tail = job->hasOut ? job->minPunchF : 0;
  for (int f = 0; f < nf; f++) {
    int64_t nv = (db->fileStart[f + 1] - db->fileStart[f]) - tail - job->qin.W + 1;
    out[f].numOffsets = (int32_t)std::max<int64_t>(nv, 0);
    out[f].maxSim = keys[f] ? float_from_order_key((uint32_t)(keys[f] >> 32)) : -INFINITY;
    out[f].maxSimOut = keysOut[f] ? float_from_order_key((uint32_t)(keysOut[f] >> 32)) : -INFINITY;
    out[f]._pad = 0;
  }
  return SGZ_OK;
}

int sgz_corr_set_global(sgz_corr *job, const sgz_file_summary *all, int32_t nFilesGlobal, int32_t myFirstFile) {
  SGZ_REQUIRE(job && (all || nFilesGlobal == 0), "NULL argument");
  if (!job->scanned) { set_error("sgz_corr_set_global before scan"); return SGZ_ERR_STATE; }
  SGZ_REQUIRE(myFirstFile >= 0 && myFirstFile + job->db->numFiles() <= nFilesGlobal,
              "local shard [%d,+%d) outside the global file list of %d", myFirstFile, job->db->numFiles(),
              nFilesGlobal);
  job->globalSummary.assign(all, all + nFilesGlobal);
  job->nFilesGlobal = nFilesGlobal;
  job->myFirst = myFirstFile;
  job->globalSet = true;
  job->allPrio.clear();
  job->nextFile = 0;
  job->finished = nFilesGlobal == 0 || job->cfg.numMatches <= 0 || job->cfg.numPerFile <= 0;
  return SGZ_OK;
}

int sgz_corr_local_top(sgz_corr *job, sgz_file_entry *out, int32_t cap, int32_t *n, int32_t *numFiles) {
  SGZ_REQUIRE(job && n && numFiles, "NULL argument");
  if (!job->scanned) { set_error("sgz_corr_local_top before scan"); return SGZ_ERR_STATE; }
  SGZ_REQUIRE(!job->hasOut, "sgz_corr_local_top is for punch-in-only searches (the punch-out merge needs every file's length)");
  sgz_db *db = job->db;
  const int nf = db->numFiles();
  *numFiles = nf;
  const int k = std::min(nf, std::max(job->cfg.numMatches, 0));
  *n = k;
  if (!out) return SGZ_OK;
  SGZ_REQUIRE(cap >= k, "entry buffer too small (%d < %d)", cap, k);
  job->localSummary.resize((size_t)std::max(nf, 1));
  int32_t got = 0;
  SGZ_TRY(sgz_corr_local_summary(job, job->localSummary.data(), nf, &got));
  std::vector<int32_t> idx((size_t)nf);
  for (int f = 0; f < nf; f++) idx[(size_t)f] = f;
  auto greater = [&](int32_t a, int32_t b) {
    const float x = job->localSummary[(size_t)a].maxSim, y = job->localSummary[(size_t)b].maxSim;
    return x != y ? x > y : a < b;
  };
  if (k < nf) std::nth_element(idx.begin(), idx.begin() + k, idx.end(), greater);
  std::sort(idx.begin(), idx.begin() + k);
  for (int i = 0; i < k; i++) out[i] = sgz_file_entry{idx[(size_t)i], job->localSummary[(size_t)idx[(size_t)i]].maxSim};
  return SGZ_OK;
}

int sgz_corr_local_best(sgz_corr *job, sgz_record *out, int32_t cap, int32_t *n, int32_t *numFiles, int32_t *ok) {
  SGZ_REQUIRE(job && n && numFiles && ok, "NULL argument");
  if (!job->scanned) { set_error("sgz_corr_local_best before scan"); return SGZ_ERR_STATE; }
  sgz_db *db = job->db;
  const int nf = db->numFiles();
  *numFiles = nf;
  const int k = std::min(nf, std::max(job->cfg.numMatches, 0));
  *n = k;
  *ok = 0;
  if (!out) return SGZ_OK;
  SGZ_REQUIRE(cap >= k, "record buffer too small (%d < %d)", cap, k);
  job->localSummary.resize((size_t)std::max(nf, 1));
  int32_t got = 0;
  SGZ_TRY(sgz_corr_local_summary(job, job->localSummary.data(), nf, &got));   // (fetches keys, NaN flags and boosts)
  const bool direct = job->direct && (nf == 0 || db->usedFrames == 0 || (job->keysCached && (int)job->hKeys.size() == nf));
  bool anyNaN = false;
  if (direct && job->keysCached)
    for (int f = 0; f < nf; f++) anyNaN |= job->hNaN[(size_t)f] != 0;
  *ok = direct && !anyNaN;
  std::vector<int32_t> idx((size_t)nf);
  for (int f = 0; f < nf; f++) idx[(size_t)f] = f;
  auto greater = [&](int32_t a, int32_t b) {
    const float x = job->localSummary[(size_t)a].maxSim, y = job->localSummary[(size_t)b].maxSim;
    return x != y ? x > y : a < b;
  };
  if (k < nf) std::nth_element(idx.begin(), idx.begin() + k, idx.end(), greater);
  std::sort(idx.begin(), idx.begin() + k);
  for (int i = 0; i < k; i++) {
    const int f = idx[(size_t)i];
    sgz_record r{f, 0, -1, -1, job->localSummary[(size_t)f].maxSim, 1.0f, 1.0f, 0};
    if (*ok && job->keysCached && job->hKeys[(size_t)f] != 0ull) {
      r.kind = 1;                                            // the file has an entry
      r.piOff = (int32_t)(0xffffffffu - (uint32_t)job->hKeys[(size_t)f]);
      r.boostIn = job->hBoost[(size_t)f];
    }
    out[i] = r;
  }
  return SGZ_OK;
}

int sgz_corr_finish_from_best(sgz_corr *job, const sgz_record *all, int32_t nAll, int32_t nFilesGlobal) {
  SGZ_REQUIRE(job && (all || nAll == 0), "NULL argument");
  if (!job->scanned) { set_error("sgz_corr_finish_from_best before scan"); return SGZ_ERR_STATE; }
  SGZ_REQUIRE(!job->hasOut && job->cfg.numPerFile == 1, "sgz_corr_finish_from_best is for punch-in searches with numPerFile = 1");
  const int K = job->cfg.numMatches, step = job->step, W = job->qin.W;
  std::vector<sgz_record> recs(all, all + nAll);
  std::stable_sort(recs.begin(), recs.end(), [](const sgz_record &a, const sgz_record &b) { return a.file < b.file; });
  job->allPrio.clear();
  if (K > 0)
    for (const sgz_record &r : recs) {                       // the files in list order, FeatureCorrelationImpl.scala:160-165
      SGZ_REQUIRE(r.file >= 0 && r.file < nFilesGlobal, "record names file %d of %d", r.file, nFilesGlobal);
      if (r.kind != 1) continue;                             // no valid offset in that file
      // entryHasSpace || sim > lowestSim (:120-129) with an empty entryPrio, then allPrio ++= entryPrio; take(numMatches)
      if ((int)job->allPrio.size() < K || r.sim > job->allPrio.back().sim) {
        sgz_match m;
        m.sim = r.sim; m.file = r.file;
        m.start = feat_to_full(r.piOff, step);
        m.stop = feat_to_full(r.piOff + W, step);
        m.boostIn = r.boostIn; m.boostOut = 1.0f;
        allprio_add(job->allPrio, m);
        if ((int)job->allPrio.size() > K) job->allPrio.resize(K);
      }
    }
  job->nFilesGlobal = nFilesGlobal;
  job->nextFile = nFilesGlobal;
  job->globalSet = true;
  job->finished = true;
  job->progress = 1.0f;
  return SGZ_OK;
}

int sgz_corr_set_global_top(sgz_corr *job, const sgz_file_entry *all, int32_t nAll, int32_t nFilesGlobal, int32_t myFirstFile) {
  SGZ_REQUIRE(job && (all || nAll == 0), "NULL argument");
  if (!job->scanned) { set_error("sgz_corr_set_global_top before scan"); return SGZ_ERR_STATE; }
  const int nf = job->db->numFiles();
  SGZ_REQUIRE((int)job->localSummary.size() >= nf, "sgz_corr_set_global_top before sgz_corr_local_top");
  SGZ_REQUIRE(myFirstFile >= 0 && myFirstFile + nf <= nFilesGlobal, "local shard [%d,+%d) outside the global file list of %d",
              myFirstFile, nf, nFilesGlobal);
  std::vector<sgz_file_summary> g((size_t)nFilesGlobal, sgz_file_summary{-INFINITY, 0, -INFINITY, 0});
  for (int i = 0; i < nAll; i++) {
    SGZ_REQUIRE(all[i].file >= 0 && all[i].file < nFilesGlobal, "entry %d names file %d of %d", i, all[i].file, nFilesGlobal);
    g[(size_t)all[i].file].maxSim = all[i].maxSim;
  }
  for (int f = 0; f < nf; f++) g[(size_t)(myFirstFile + f)] = job->localSummary[(size_t)f];   // own files: exact
  return sgz_corr_set_global(job, g.data(), nFilesGlobal, myFirstFile);
}

// One selection round on the local shard.  See the protocol in strugatzki_b200.h.
int sgz_corr_select(sgz_corr *job, int32_t *nRecords) {
  SGZ_REQUIRE(job && nRecords, "NULL argument");
  if (!job->globalSet) { set_error("sgz_corr_select before set_global"); return SGZ_ERR_STATE; }
  if (job->abortFlag) return SGZ_ERR_ABORTED;
  sgz_ctx *ctx = job->ctx;
  sgz_db *db = job->db;
  SGZ_TRY(ctx->bind());
  job->localRecords.clear();
  *nRecords = 0;
  job->summaryPrefetched = false;    // the selection rounds reuse the pinned scratch
  if (job->finished) return SGZ_OK;
  if (job->hasOut) return corr_select_punchout(job, nRecords);
  const int K = job->cfg.numMatches, npf = job->cfg.numPerFile;
  const int room = K - (int)job->allPrio.size();
  const int myLo = job->myFirst, myHi = job->myFirst + db->numFiles();
  const int tail = 0;
  // numPerFile = 1 on the tensor-core scan: the entry of a file is its maximum at its first position (addMatch keeps one
  // match per file and replaces it only by a strictly larger sim, FeatureCorrelationImpl.scala:135-150), which the scan left
  // in fileMax together with its boost -- no replay, no candidate kernel, no device round trip.  Exception: a file with a NaN
  // window in a FILLING round (NaN enters an entry that still has space and then blocks it, :120-129) keeps the replay.
  const bool direct = job->direct && job->keysCached && (int)job->hKeys.size() == db->numFiles();
  auto direct_record = [&](int f) {
    const unsigned long long key = job->hKeys[(size_t)f];
    if (key == 0ull) return;                               // no valid offset in this file
    const sgz_record r{myLo + f, 1, (int32_t)(0xffffffffu - (uint32_t)key), -1, float_from_order_key((uint32_t)(key >> 32)),
                       job->hBoost[(size_t)f], 1.0f, 0};
    job->localRecords.push_back(r);
  };
  if (room > 0) {
    // ---- filling round: files whose maxEntrySz is known without looking at their results ----
    const int nb = room >= npf ? room / npf : 1;
    const int m = room >= npf ? npf : room;
    job->roundKind = 0;
    job->roundFirst = job->nextFile;
    job->roundCount = std::min(nb, job->nFilesGlobal - job->nextFile);
    job->roundMaxEntrySz = m;
    std::vector<int32_t> files, directFiles;
    for (int g = job->roundFirst; g < job->roundFirst + job->roundCount; g++)
      if (g >= myLo && g < myHi) {
        if (direct && !job->hNaN[(size_t)(g - myLo)]) directFiles.push_back(g - myLo);
        else files.push_back(g - myLo);
      }
    for (int f : directFiles) direct_record(f);
    if (!files.empty()) {
      const int nj = (int)files.size();
      SGZ_TRY(job->dFiles.alloc(nj));
      SGZ_TRY(job->dCounts.alloc(nj));
      SGZ_TRY(job->dEntries.alloc((size_t)nj * (npf + 1)));
      SGZ_CUDA(cudaMemcpyAsync(job->dFiles.p, files.data(), nj * sizeof(int32_t), cudaMemcpyHostToDevice,
                               ctx->stream));
      FillParams fp{};
      fp.sim = job->simIn.p; fp.boost = boost_src(job, job->qin, job->boostIn.p); fp.fileStart = db->dFileStart.p;
      fp.files = job->dFiles.p; fp.numJobs = nj; fp.W = job->qin.W; fp.tailExtra = tail;
      fp.numPerFile = npf; fp.maxEntrySz = m; fp.minSpacing = job->cfg.minSpacing; fp.step = job->step;
      fp.entries = job->dEntries.p; fp.counts = job->dCounts.p;
      SGZ_TRY(ctx->begin_call());
      k_replay_fill<<<(unsigned)nj, kFillPiThreads, 0, ctx->stream>>>(fp);
      SGZ_LAUNCH_CHECK(ctx);
      SGZ_TRY(ctx->end_call_async());
      const size_t entBytes = (size_t)nj * (npf + 1) * sizeof(EntryRec);
      SGZ_TRY(job->pin(entBytes + (size_t)nj * sizeof(int32_t)));
      const EntryRec *ents = reinterpret_cast<const EntryRec *>(job->hPin);
      const int32_t *counts = reinterpret_cast<const int32_t *>(job->hPin + entBytes);
      SGZ_CUDA(cudaMemcpyAsync(job->hPin, job->dEntries.p, entBytes, cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(job->hPin + entBytes, job->dCounts.p, nj * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
      SGZ_TRY(ctx->collect_call());
      job->selectMs += ctx->lastMs;
      for (int j = 0; j < nj; j++)
        for (int k = 0; k < counts[j]; k++) {
          const EntryRec &e = ents[(size_t)j * (npf + 1) + k];
          sgz_record r{myLo + files[j], 1, e.piOff, -1, e.sim, e.boostIn, e.boostOut, 0};
          job->localRecords.push_back(r);
        }
    }
  } else {
    // ---- full round: allPrio is full; candidates above a provable lower bound of allPrio.last ----
    job->roundKind = 1;
    job->roundFirst = job->nextFile;
    job->roundCount = job->nFilesGlobal - job->nextFile;
    // running K-th largest DISTINCT value of (allPrio sims U file maxima so far): ascending vector
    std::vector<float> top;
    for (const sgz_match &mm : job->allPrio)
      if (mm.sim == mm.sim) top.push_back(mm.sim);
    std::sort(top.begin(), top.end());
    top.erase(std::unique(top.begin(), top.end()), top.end());
    const float theta0 = job->allPrio.back().sim;   // may be NaN only if every element is NaN
    std::vector<int32_t> files;
    std::vector<float> thr;
    for (int g = job->roundFirst; g < job->nFilesGlobal; g++) {
      float bound = theta0;
      if ((int)top.size() >= K && top[top.size() - K] > bound) bound = top[top.size() - K];
      if (!(bound == bound)) break;  // allPrio.last is NaN: `sim > NaN` never holds -> nothing more is accepted
      const float mx = job->globalSummary[g].maxSim;
      if (mx > bound) {
        if (g >= myLo && g < myHi) {
          if (direct) direct_record(g - myLo);             // (a NaN never passes `sim > bound`: the maximum is all it takes)
          else { files.push_back(g - myLo); thr.push_back(bound); }
        }
        auto it = std::lower_bound(top.begin(), top.end(), mx);
        if (it == top.end() || *it != mx) top.insert(it, mx);
        if ((int)top.size() > K) top.erase(top.begin());
      }
    }
    if (!files.empty()) {
      const int nj = (int)files.size();
      SGZ_TRY(job->dFiles.alloc(nj));
      SGZ_TRY(job->dThr.alloc(nj));
      SGZ_TRY(job->dCounter.alloc(1));
      SGZ_CUDA(cudaMemcpyAsync(job->dFiles.p, files.data(), nj * sizeof(int32_t), cudaMemcpyHostToDevice,
                               ctx->stream));
      SGZ_CUDA(cudaMemcpyAsync(job->dThr.p, thr.data(), nj * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      int cap = std::max(1 << 14, (int)job->dRecs.n);
      for (;;) {
        SGZ_TRY(job->dRecs.alloc(cap));
        SGZ_CUDA(cudaMemsetAsync(job->dCounter.p, 0, sizeof(int), ctx->stream));
        CandParams cp{};
        cp.sim = job->simIn.p; cp.boost = boost_src(job, job->qin, job->boostIn.p); cp.fileStart = db->dFileStart.p;
        cp.files = job->dFiles.p; cp.thresholds = job->dThr.p; cp.numJobs = nj; cp.W = job->qin.W;
        cp.tailExtra = tail; cp.fileBase = myLo; cp.out = job->dRecs.p; cp.cap = cap; cp.counter = job->dCounter.p;
        SGZ_TRY(ctx->begin_call());
        k_candidates<<<dim3((unsigned)nj, 8), 256, 0, ctx->stream>>>(cp);
        SGZ_LAUNCH_CHECK(ctx);
        SGZ_TRY(ctx->end_call_async());
        // the counter and the first records in one round trip; more records (rare) in a second one
        const int first = std::min(cap, 2048);
        SGZ_TRY(job->pin(16 + (size_t)first * sizeof(sgz_record)));
        SGZ_CUDA(cudaMemcpyAsync(job->hPin, job->dCounter.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SGZ_CUDA(cudaMemcpyAsync(job->hPin + 16, job->dRecs.p, (size_t)first * sizeof(sgz_record), cudaMemcpyDeviceToHost,
                                 ctx->stream));
        SGZ_CUDA(cudaStreamSynchronize(ctx->stream));
        SGZ_TRY(ctx->collect_call());
        job->selectMs += ctx->lastMs;
        const int count = *reinterpret_cast<const int *>(job->hPin);
        if (count <= cap) {
          job->localRecords.resize(count);
          if (count > 0) memcpy(job->localRecords.data(), job->hPin + 16, (size_t)std::min(count, first) * sizeof(sgz_record));
          if (count > first) {
            SGZ_CUDA(cudaMemcpyAsync(job->localRecords.data() + first, job->dRecs.p + first,
                                     (size_t)(count - first) * sizeof(sgz_record), cudaMemcpyDeviceToHost, ctx->stream));
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

int sgz_corr_records(sgz_corr *job, sgz_record *out, int32_t cap, int32_t *n) {
  SGZ_REQUIRE(job && n, "NULL argument");
  *n = (int32_t)job->localRecords.size();
  if (!out) return SGZ_OK;
  SGZ_REQUIRE(cap >= *n, "record buffer too small (%d < %d)", cap, *n);
  if (*n > 0) memcpy(out, job->localRecords.data(), (size_t)*n * sizeof(sgz_record));
  return SGZ_OK;
}

int sgz_corr_merge(sgz_corr *job, const sgz_record *all, int32_t nAll, int32_t *done) {
  SGZ_REQUIRE(job && done && (all || nAll == 0), "NULL argument");
  if (!job->globalSet) { set_error("sgz_corr_merge before set_global"); return SGZ_ERR_STATE; }
  if (job->finished) { *done = 1; return SGZ_OK; }
  if (job->hasOut) return corr_merge_punchout(job, all, nAll, done);
  const int K = job->cfg.numMatches, npf = job->cfg.numPerFile, step = job->step, W = job->qin.W;
  std::vector<sgz_record> recs(all, all + nAll);
  auto to_match = [&](const EntryRec &e, int file) {
    sgz_match m;
    m.sim = e.sim; m.file = file;
    m.start = feat_to_full(e.piOff, step);
    m.stop = feat_to_full(e.stopOff, step);
    m.boostIn = e.boostIn; m.boostOut = e.boostOut;
    return m;
  };
  auto merge_entry = [&](const EntryRec *e, int n, int file) {  // allPrio ++= entryPrio; take(numMatches)
    for (int i = 0; i < n; i++) allprio_add(job->allPrio, to_match(e[i], file));
    if ((int)job->allPrio.size() > K) job->allPrio.resize(K);
  };
  if (job->roundKind == 0) {
    std::stable_sort(recs.begin(), recs.end(), [](const sgz_record &a, const sgz_record &b) { return a.file < b.file; });
    size_t i = 0;
    std::vector<EntryRec> ent;
    for (int g = job->roundFirst; g < job->roundFirst + job->roundCount; g++) {
      ent.clear();
      while (i < recs.size() && recs[i].file < g) i++;
      while (i < recs.size() && recs[i].file == g) {
        if (recs[i].kind == 1) ent.push_back(EntryRec{recs[i].sim, recs[i].piOff, recs[i].piOff + W, recs[i].boostIn, recs[i].boostOut});
        i++;
      }
      merge_entry(ent.data(), (int)ent.size(), g);
    }
    job->nextFile = job->roundFirst + job->roundCount;
  } else {
    std::sort(recs.begin(), recs.end(), [](const sgz_record &a, const sgz_record &b) {
      return a.file != b.file ? a.file < b.file : a.piOff < b.piOff;
    });
    std::vector<EntryRec> store((size_t)npf + 1);
    size_t i = 0;
    while (i < recs.size()) {
      const int g = recs[i].file;
      Machine mc;
      const int allSize = (int)job->allPrio.size();
      mc.reset(store.data(), npf, std::min(K - allSize, npf), allSize > 0, allSize > 0 ? job->allPrio.back().sim : 0.f,
               job->cfg.minSpacing, step);
      for (; i < recs.size() && recs[i].file == g; i++) {
        const sgz_record &r = recs[i];
        if (mc.has_space() || r.sim > mc.lowest())
          mc.add(EntryRec{r.sim, r.piOff, r.piOff + W, r.boostIn, r.boostOut});
      }
      merge_entry(mc.e, mc.n, g);
    }
    job->nextFile = job->nFilesGlobal;
  }
  if (job->nextFile >= job->nFilesGlobal) job->finished = true;
  *done = job->finished ? 1 : 0;
  if (job->finished) job->progress = 1.0f;
  return SGZ_OK;
}

static int corr_run_body(sgz_corr *job) {
  SGZ_TRY(sgz_corr_scan(job));
  int32_t nf = 0;
  SGZ_TRY(sgz_corr_local_summary(job, nullptr, 0, &nf));
  std::vector<sgz_file_summary> sum((size_t)std::max(nf, 1));
  SGZ_TRY(sgz_corr_local_summary(job, sum.data(), nf, &nf));
  SGZ_TRY(sgz_corr_set_global(job, sum.data(), nf, 0));
  int32_t done = job->finished ? 1 : 0;
  while (!done) {
    if (job->abortFlag) return SGZ_ERR_ABORTED;
    int32_t nrec = 0;
    SGZ_TRY(sgz_corr_select(job, &nrec));
    SGZ_TRY(sgz_corr_merge(job, job->localRecords.data(), nrec, &done));
    if (job->nFilesGlobal > 0)
      job->progress = 0.8f + 0.2f * (float)job->nextFile / (float)job->nFilesGlobal;
  }
  job->progress = 1.0f;
  return SGZ_OK;
}

int sgz_corr_run(sgz_corr *job) {
  SGZ_REQUIRE(job, "job is NULL");
  job->abortFlag = 0;
  return corr_run_body(job);
}

int sgz_corr_start(sgz_corr *job) {
  SGZ_REQUIRE(job, "job is NULL");
  if (job->worker.joinable()) {
    if (!job->doneFlag) { set_error("job already running"); return SGZ_ERR_STATE; }
    job->worker.join();
  }
  job->abortFlag = 0;
  job->doneFlag = 0;
  job->status = 0;
  job->progress = 0.f;
  job->worker = std::thread([job]() {
    int rc = corr_run_body(job);
    job->status = rc;
    job->doneFlag = 1;
  });
  return SGZ_OK;
}

int sgz_corr_poll(sgz_corr *job, float *progress, int32_t *done, int32_t *status) {
  SGZ_REQUIRE(job, "job is NULL");
  if (progress) *progress = job->progress.load();
  if (done) *done = job->doneFlag.load();
  if (status) *status = job->status.load();
  return SGZ_OK;
}

int sgz_corr_abort(sgz_corr *job) {
  SGZ_REQUIRE(job, "job is NULL");
  job->abortFlag = 1;
  return SGZ_OK;
}

int sgz_corr_wait(sgz_corr *job) {
  SGZ_REQUIRE(job, "job is NULL");
  if (job->worker.joinable()) job->worker.join();
  return job->status.load();
}

int sgz_corr_result(sgz_corr *job, sgz_match *out, int32_t cap, int32_t *n) {
  SGZ_REQUIRE(job && n, "NULL argument");
  *n = (int32_t)job->allPrio.size();
  if (!out) return SGZ_OK;
  SGZ_REQUIRE(cap >= *n, "result buffer too small (%d < %d)", cap, *n);
  for (int i = 0; i < *n; i++) out[i] = job->allPrio[i];
  return SGZ_OK;
}

int sgz_corr_num_offsets(sgz_corr *job, int64_t *n) {
  SGZ_REQUIRE(job && n, "NULL argument");
  *n = job->numOffsets;
  return SGZ_OK;
}

int sgz_corr_timing(sgz_corr *job, float *scanMs, float *selectMs, int64_t *scanLaunches) {
  SGZ_REQUIRE(job, "job is NULL");
  if (scanMs) *scanMs = job->scanMs;
  if (selectMs) *selectMs = job->selectMs;
  if (scanLaunches) *scanLaunches = job->scanLaunches;
  return SGZ_OK;
}

int sgz_corr_curve(sgz_corr *job, int32_t which, int32_t file, int64_t first, int64_t n, float *sim, float *boost) {
  SGZ_REQUIRE(job, "job is NULL");
  if (!job->scanned) { set_error("sgz_corr_curve before scan"); return SGZ_ERR_STATE; }
  sgz_db *db = job->db;
  SGZ_REQUIRE(which == 0 || (which == 1 && job->hasOut), "curve %d not available", which);
  SGZ_REQUIRE(file >= 0 && file < db->numFiles(), "file index %d out of range", file);
  const int W = which == 0 ? job->qin.W : job->qout.W;
  const int tail = (which == 0 && job->hasOut) ? job->minPunchF : 0;
  const int64_t nValid = (db->fileStart[file + 1] - db->fileStart[file]) - tail - W + 1;
  SGZ_REQUIRE(first >= 0 && n >= 0 && first + n <= std::max<int64_t>(nValid, 0),
              "curve range [%lld,+%lld) outside the %lld evaluated offsets", (long long)first, (long long)n,
              (long long)nValid);
  SGZ_TRY(job->ctx->bind());
  const float *s = (which == 0 ? job->simIn.p : job->simOut.p) + db->fileStart[file] + first;
  const float *bArr = which == 0 ? job->boostIn.p : job->boostOut.p;
  if (n > 0) {
    if (sim) SGZ_CUDA(cudaMemcpyAsync(sim, s, n * sizeof(float), cudaMemcpyDeviceToHost, job->ctx->stream));
    if (boost && bArr) {
      SGZ_CUDA(cudaMemcpyAsync(boost, bArr + db->fileStart[file] + first, n * sizeof(float), cudaMemcpyDeviceToHost,
                               job->ctx->stream));
    } else if (boost) {   // tensor-core scan: the boost values are computed on demand
      DevBuf<float> tmp;
      SGZ_TRY(tmp.alloc((size_t)n));
      k_boost_curve<<<(unsigned)ceil_div<int64_t>(n, 128), 128, 0, job->ctx->stream>>>(
          boost_src(job, which == 0 ? job->qin : job->qout, nullptr), db->fileStart[file] + first, first, n, tmp.p);
      SGZ_LAUNCH_CHECK(job->ctx);
      SGZ_CUDA(cudaMemcpyAsync(boost, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost, job->ctx->stream));
      SGZ_CUDA(cudaStreamSynchronize(job->ctx->stream));
    }
    SGZ_CUDA(cudaStreamSynchronize(job->ctx->stream));
  }
  return SGZ_OK;
}

}  // extern "C"

#include "api_segself.cuh"
