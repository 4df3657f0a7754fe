// db.cuh -- the feature database resident in HBM.
//
// Replaces the reference's per-offset feature I/O: `afExtr.read(eInBuf, readOff, 1)` +
// `MathUtil.normalize` once per frame-offset (FeatureCorrelationImpl.scala:195-197,
// MathUtil.scala:132-152).  Every DB file is uploaded once, normalised with the same float
// expression `(f - min) / (max - min)` (IEEE division, bit-identical to the JVM), and kept as ONE
// stream in which the files follow each other without padding.
//
// HBM layout ("pair rows"): data[pair p][global frame g] is a float2 holding channels (2p, 2p+1);
// an odd channel count is padded with a zero channel.  K1 streams one pair row per TMA bulk copy and
// feeds it to packed FFMA2 (fma.rn.f32x2) whose operands are exactly these naturally aligned
// pairs -- measured on B200, scalar FFMA with live register operands tops out near 51 TFLOP/s while
// FFMA2 reaches the full 73 TFLOP/s (tools/peaks_probe.py).
#pragma once
#include "common.cuh"

struct sgz_db {
  std::atomic<int> refs{0};             // jobs created on this database
  bool zombie = false;                  // sgz_db_destroy called while jobs were alive
  sgz_ctx *ctx = nullptr;
  int numCh = 0;
  int numPairs = 0;
  bool hasNorm = false;
  std::vector<float> norm;              // [numCh][2] host copy
  DevBuf<float> dNorm;                  // [numCh][2]; {0,1} per channel when normalize=false
  DevBuf<float2> dData;                 // pair rows [numPairs][capFrames]
  int64_t capFrames = 0;                // row stride in frames (multiple of 1024)
  int64_t usedFrames = 0;
  std::vector<int64_t> fileStart;       // size numFiles+1 (last = usedFrames)
  DevBuf<int64_t> dFileStart;
  bool finalized = false;
  // upload staging ring: H2D copies on copyStream run ahead of the prepare kernels on ctx->stream.  The ring is
  // deep enough to keep PCIe busy while prepare kernels queue behind a K1 launch of a streaming scan.
  static constexpr int kStageSlots = 32;
  DevBuf<unsigned char> dStage[kStageSlots];
  cudaStream_t copyStream = nullptr;
  cudaEvent_t stageFull[kStageSlots] = {}, stageFree[kStageSlots] = {};
  int stageIdx = 0;
  bool stageUsed[kStageSlots] = {};
  // upload progress markers on ctx->stream: every frame below `uptoFrame` is resident once `ev` has fired.
  // A scan that finds markers (sgz_db_finalize_async) launches K1 range by range behind them.
  struct Chunk { int64_t uptoFrame; cudaEvent_t ev; };
  std::vector<Chunk> chunks;
  int64_t chunkMark = 0;
  // copy coalescing: HOST_STABLE files whose host buffers follow each other in memory (one arena, an mmap-ed database
  // cache) are uploaded as ONE copy of up to kRunBytes -- file-sized copies (2.9 MB) reach 51 of the 55.5 GB/s PCIe gives
  // a large copy (profiles/r01_pcie_probe.json).  The run is flushed by the next file that does not continue it and by
  // everything that needs the frames on the device.
  struct Pend { size_t off; int64_t nFrames; int layout; int64_t dst; };
  static constexpr size_t kRunBytes = (size_t)32 << 20;
  const unsigned char *pendPtr = nullptr;
  size_t pendBytes = 0;
  std::vector<Pend> pend;

  // tensor-core K1 (corr_tc2.cuh): pre-swizzled FP16 planes [numCh * 2][planeStrideBytes] and the tile-transposed
  // per-frame sums, built from the pair rows on demand (frames below planesUpto are done; a patch resets it)
  DevBuf<unsigned char> dPlanes;
  DevBuf<uint2> dSideA;                 // per frame (loudness, spectral sum) as Double high words, tile transposed
  DevBuf<uint32_t> dSideB;              // per frame spectral sum of squares
  DevBuf<double> dB16;                  // [16][planeRows] FP64 sums of aligned 16-frame blocks
  int64_t planeStrideBytes = 0, planesUpto = 0, planeRows = 0;

  int numFiles() const { return (int)fileStart.size() - 1; }
};

namespace sgz {

constexpr int64_t kDbSlack = 16384;  // readable zero frames behind the last file (tile halo)
// upload marker granularity (~80 files of 10 min, 4 ms of PCIe); SGZ_CHUNK_FRAMES overrides it (tests)
inline int64_t chunk_frames() {
  static const int64_t v = [] {
    const char *e = getenv("SGZ_CHUNK_FRAMES");
    return e && atoll(e) > 0 ? (int64_t)atoll(e) : (int64_t)(4 << 20);
  }();
  return v;
}

// ---- synthetic features.  Integer hash -> exact integer sum of 8
// consecutive 24-bit values -> ONE float scale, so numpy (strugatzki_b200/synth.py) and this
// kernel produce bit-identical floats. --------------------------------------------------------
__host__ __device__ inline uint32_t synth_u24(uint64_t seed, uint32_t stream, uint32_t c, uint64_t t) {
  uint64_t z = seed + (uint64_t)stream * 0xD1B54A32D192ED03ULL + (uint64_t)c * 0x9E3779B97F4A7C15ULL +
               t * 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (uint32_t)(z >> 40);
}

__device__ inline float synth_value(uint64_t seed, uint32_t stream, uint32_t c, uint64_t t, float mu,
                                    float sigma) {
  int32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += (int32_t)synth_u24(seed, stream, c, t + k);
  float g = __fmul_rn((float)(s - 67108860), __int_as_float(0x339cc471));  // * sqrt(1.5)/2^24
  return __fadd_rn(mu, __fmul_rn(sigma, g));
}

__device__ inline float normalize_value(float f, float mn, float mx) {
  return __fdiv_rn(__fsub_rn(f, mn), __fsub_rn(mx, mn));  // MathUtil.scala:140-147
}

// one thread = one frame of one channel pair -> coalesced float2 stores
__global__ void k_db_synth(float2 *__restrict__ data, int64_t rowStride, int64_t dstFrame, int64_t nFrames,
                           int numCh, uint64_t seed, uint32_t stream, const float *__restrict__ mu,
                           const float *__restrict__ sigma, float floor0, const float *__restrict__ norm) {
  const int p = blockIdx.y, c0 = 2 * p, c1 = 2 * p + 1;
  const bool has1 = c1 < numCh;
  const float m0 = mu[c0], s0 = sigma[c0], mn0 = norm[2 * c0], mx0 = norm[2 * c0 + 1];
  const float m1 = has1 ? mu[c1] : 0.f, s1 = has1 ? sigma[c1] : 0.f;
  const float mn1 = has1 ? norm[2 * c1] : 0.f, mx1 = has1 ? norm[2 * c1 + 1] : 1.f;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nFrames;
       t += (int64_t)gridDim.x * blockDim.x) {
    float x = synth_value(seed, stream, (uint32_t)c0, (uint64_t)t, m0, s0);
    if (c0 == 0) x = fmaxf(x, floor0);
    float2 v;
    v.x = normalize_value(x, mn0, mx0);
    v.y = has1 ? normalize_value(synth_value(seed, stream, (uint32_t)c1, (uint64_t)t, m1, s1), mn1, mx1) : 0.f;
    data[(int64_t)p * rowStride + dstFrame + t] = v;
  }
}

// the same for `numFiles` equally long files in ONE launch (streams stream0, stream0 + 1, ...): the 1000 h bench
// database is 6000 files, and 6000 separate launches made the ncu launch list uncapturable
__global__ void k_db_synth_many(float2 *__restrict__ data, int64_t rowStride, int64_t dstFrame, int64_t nFrames,
                                int64_t numFiles, int numCh, uint64_t seed, uint32_t stream0,
                                const float *__restrict__ mu, const float *__restrict__ sigma, float floor0,
                                const float *__restrict__ norm) {
  const int p = blockIdx.y, c0 = 2 * p, c1 = 2 * p + 1;
  const bool has1 = c1 < numCh;
  const float m0 = mu[c0], s0 = sigma[c0], mn0 = norm[2 * c0], mx0 = norm[2 * c0 + 1];
  const float m1 = has1 ? mu[c1] : 0.f, s1 = has1 ? sigma[c1] : 0.f;
  const float mn1 = has1 ? norm[2 * c1] : 0.f, mx1 = has1 ? norm[2 * c1 + 1] : 1.f;
  const int64_t total = nFrames * numFiles;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i / nFrames, t = i - f * nFrames;
    const uint32_t stream = stream0 + (uint32_t)f;
    float x = synth_value(seed, stream, (uint32_t)c0, (uint64_t)t, m0, s0);
    if (c0 == 0) x = fmaxf(x, floor0);
    float2 v;
    v.x = normalize_value(x, mn0, mx0);
    v.y = has1 ? normalize_value(synth_value(seed, stream, (uint32_t)c1, (uint64_t)t, m1, s1), mn1, mx1) : 0.f;
    data[(int64_t)p * rowStride + dstFrame + i] = v;
  }
}

__device__ inline float load_be(const float *p) {
  uint32_t v = *reinterpret_cast<const uint32_t *>(p);
  return __int_as_float((int)__byte_perm(v, 0, 0x0123));
}

// destination index of (channel c, frame t): pair rows (DB) or plain planar rows (segm / selfsim)
__device__ __forceinline__ int64_t dst_index(int pairRows, int c, int64_t t, int64_t rowStride) {
  return pairRows ? ((int64_t)(c >> 1) * rowStride + t) * 2 + (c & 1) : (int64_t)c * rowStride + t;
}

// K0: interleaved (LE or BE) / planar staging -> normalised rows.  One block converts kPrepFrames
// frames; the interleaved source is read with fully coalesced 4-byte loads into shared memory.
constexpr int kPrepFrames = 256;
__global__ void k_db_prepare(const float *__restrict__ src, int layout, int64_t srcFrames,
                             float *__restrict__ data, int64_t rowStride, int64_t dstFrame, int numCh,
                             const float *__restrict__ norm, int pairRows) {
  extern __shared__ float sh[];  // [kPrepFrames * numCh]
  int64_t f0 = (int64_t)blockIdx.x * kPrepFrames;
  int nf = (int)min((int64_t)kPrepFrames, srcFrames - f0);
  if (nf <= 0) return;
  if (layout == SGZ_LAYOUT_PLANAR_LE) {
    for (int i = threadIdx.x; i < nf * numCh; i += blockDim.x) {
      int c = i / nf, t = i - c * nf;
      float v = src[(int64_t)c * srcFrames + f0 + t];
      data[dst_index(pairRows, c, dstFrame + f0 + t, rowStride)] = normalize_value(v, norm[2 * c], norm[2 * c + 1]);
    }
    return;
  }
  const float *s = src + f0 * numCh;
  for (int i = threadIdx.x; i < nf * numCh; i += blockDim.x)
    sh[i] = layout == SGZ_LAYOUT_INTERLEAVED_BE ? load_be(s + i) : s[i];
  __syncthreads();
  if (pairRows) {
    const int np = (numCh + 1) >> 1;
    for (int i = threadIdx.x; i < nf * np; i += blockDim.x) {   // one float2 per thread, coalesced
      int p = i / nf, t = i - p * nf;
      int c0 = 2 * p, c1 = 2 * p + 1;
      float2 v;
      v.x = normalize_value(sh[t * numCh + c0], norm[2 * c0], norm[2 * c0 + 1]);
      v.y = c1 < numCh ? normalize_value(sh[t * numCh + c1], norm[2 * c1], norm[2 * c1 + 1]) : 0.f;
      reinterpret_cast<float2 *>(data)[(int64_t)p * rowStride + dstFrame + f0 + t] = v;
    }
  } else {
    for (int i = threadIdx.x; i < nf * numCh; i += blockDim.x) {
      int c = i / nf, t = i - c * nf;
      data[(int64_t)c * rowStride + dstFrame + f0 + t] = normalize_value(sh[t * numCh + c], norm[2 * c], norm[2 * c + 1]);
    }
  }
}

// planar [numCh][n] read-back of normalised frames from the pair rows
__global__ void k_db_gather(const float2 *__restrict__ data, int64_t rowStride, int64_t srcFrame, int64_t n,
                            int numCh, float *__restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * numCh) return;
  int c = (int)(i / n);
  int64_t t = i - (int64_t)c * n;
  float2 v = data[(int64_t)(c >> 1) * rowStride + srcFrame + t];
  out[i] = (c & 1) ? v.y : v.x;
}

inline int db_grow(sgz_db *db, int64_t needFrames) {
  if (needFrames + kDbSlack <= db->capFrames) return SGZ_OK;
  int64_t newCap = std::max<int64_t>(needFrames + kDbSlack, db->capFrames * 2);
  newCap = (newCap + 1023) / 1024 * 1024;
  DevBuf<float2> nd;
  SGZ_TRY(nd.alloc((size_t)newCap * db->numPairs));
  cudaStream_t st = db->ctx->stream;
  // no full memset: every frame below usedFrames is written by a prepare / synth kernel, and the slack
  // behind the last file is zeroed by sgz_db_finalize (pool blocks are handed out dirty)
  if (db->usedFrames > 0) {
    SGZ_CUDA(cudaMemcpy2DAsync(nd.p, (size_t)newCap * sizeof(float2), db->dData.p,
                               (size_t)db->capFrames * sizeof(float2), (size_t)db->usedFrames * sizeof(float2),
                               (size_t)db->numPairs, cudaMemcpyDeviceToDevice, st));
  }
  SGZ_CUDA(cudaStreamSynchronize(st));
  std::swap(db->dData.p, nd.p);
  std::swap(db->dData.n, nd.n);
  std::swap(db->dData.bytes, nd.bytes);
  std::swap(db->dData.dev, nd.dev);
  db->capFrames = newCap;
  return SGZ_OK;
}

inline int db_launch_prepare(sgz_db *db, const float *dSrc, int layout, int64_t nFrames, int64_t dstFrame) {
  if (nFrames <= 0) return SGZ_OK;
  int blocks = (int)ceil_div<int64_t>(nFrames, kPrepFrames);
  size_t sm = (size_t)kPrepFrames * db->numCh * sizeof(float);
  k_db_prepare<<<blocks, 256, sm, db->ctx->stream>>>(dSrc, layout, nFrames, reinterpret_cast<float *>(db->dData.p),
                                                     db->capFrames, dstFrame, db->numCh, db->dNorm.p, 1);
  SGZ_LAUNCH_CHECK(db->ctx);
  return SGZ_OK;
}

}  // namespace sgz
