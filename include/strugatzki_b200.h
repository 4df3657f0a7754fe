/*
 * strugatzki_b200.h -- C ABI of the B200-native engine for Strugatzki's feature-similarity
 * hot path (FeatureCorrelation / FeatureSegmentation / SelfSimilarity).
 *
 * This is the drop-in boundary: plain pointers and sizes, opaque handles, int return codes,
 * no exceptions, no callbacks, no torch / CUDA types.  A host (Scala via JNI/Panama, C++,
 * Python ctypes) keeps the reference's Config / Processor surface
 * (Api/FeatureCorrelation.scala:27-81,105-277, Api/FeatureSegmentation.scala:29-191,
 * Api/SelfSimilarity.scala:24-297) and replaces ONLY the processor bodies
 * (Impl/FeatureCorrelationImpl.scala:32-412, Impl/FeatureSegmentationImpl.scala:31-142,
 * Impl/SelfSimilarityImpl.scala:31-180 and the arithmetic of Impl/MathUtil.scala) by calls
 * into this library.  There is NO CPU fallback: every compute entry point fails with
 * SGZ_ERR_CUDA when no sm_100 device is usable.
 *
 * Threading: the library keeps no process-global mutable state.  One handle is driven by one
 * thread at a time; sgz_job_abort() may be called from any thread.  Several contexts (one per
 * GPU, or several per GPU) may coexist in one process.
 *
 * Units: everything that the reference expresses in SAMPLE FRAMES of the original audio
 * (spans, minPunch, maxPunch, minSpacing, corrLen, Match.start/stop, Break.pos) is passed in
 * sample frames together with stepSize = fftSize / fftOverlap; the library applies the
 * reference's fullToFeat rounding ((n + step/2) / step, FeatureCorrelationImpl.scala:38)
 * itself so that host code stays a verbatim copy of the reference's Config handling.
 */
#ifndef STRUGATZKI_B200_H
#define STRUGATZKI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGZ_ABI_VERSION 1

/* ---- return codes ---- */
#define SGZ_OK            0
#define SGZ_ERR_INVALID  (-1)  /* bad argument (reference: require / IllegalArgumentException) */
#define SGZ_ERR_CUDA     (-2)  /* CUDA runtime error or no usable sm_100 device */
#define SGZ_ERR_NOMEM    (-3)
#define SGZ_ERR_ABORTED  (-4)  /* reference: Processor.Aborted */
#define SGZ_ERR_STATE    (-5)  /* call made in the wrong state (e.g. search before finalize) */
#define SGZ_ERR_IO       (-6)  /* reference: EOFException / IOException of AudioFile */

/* ---- frame layouts accepted by sgz_db_add_file / sgz_features_upload ---- */
#define SGZ_LAYOUT_INTERLEAVED_LE 0  /* [frame][channel] float32 host order                    */
#define SGZ_LAYOUT_INTERLEAVED_BE 1  /* [frame][channel] float32 big endian = raw AIFF SSND    */
#define SGZ_LAYOUT_PLANAR_LE      2  /* [channel][frame] float32 = AudioFile.buffer layout     */
/* OR-ed into the layout of sgz_db_add_file: the (pinned) host buffer stays valid and unchanged until
 * sgz_db_finalize returns (after sgz_db_finalize_async: until the first sgz_corr_scan / sgz_corr_run on the
 * database, or a later sgz_db_finalize, returns), so the call need not wait for the upload -> H2D copies of
 * consecutive files pipeline with the prepare kernels. */
#define SGZ_LAYOUT_HOST_STABLE    0x100

typedef struct sgz_ctx  sgz_ctx;   /* one GPU + one stream */
typedef struct sgz_db   sgz_db;    /* feature database resident in HBM */
typedef struct sgz_corr sgz_corr;  /* one FeatureCorrelation search (job) */

/* FeatureCorrelation.Match, Api/FeatureCorrelation.scala:54 (file -> index in add order) */
typedef struct {
  float   sim;
  int32_t file;
  int64_t start;      /* sample frames */
  int64_t stop;
  float   boostIn;
  float   boostOut;
} sgz_match;

/* FeatureSegmentation.Break, Api/FeatureSegmentation.scala:47 */
typedef struct {
  float   sim;
  int32_t _pad;
  int64_t pos;        /* sample frames */
} sgz_break;

/* FeatureCorrelation.ConfigLike, Api/FeatureCorrelation.scala:105-158, minus the file paths
 * (the host resolves databaseFolder / metaInput and hands over frames). */
typedef struct {
  int32_t stepSize;                     /* fftSize / fftOverlap of metaInput */
  int64_t punchInStart, punchInStop;    /* punchIn.span */
  float   punchInWeight;                /* punchIn.temporalWeight */
  int32_t hasPunchOut;                  /* punchOut.isDefined */
  int64_t punchOutStart, punchOutStop;
  float   punchOutWeight;
  int64_t minPunch, maxPunch;
  float   maxBoost;
  int32_t numMatches, numPerFile;
  int64_t minSpacing;
} sgz_corr_config;

/* FeatureSegmentation.ConfigLike, Api/FeatureSegmentation.scala:71-124 */
typedef struct {
  int32_t stepSize;
  int32_t hasStart, hasStop;            /* Span.HasStart / Span.HasStop of config.span */
  int64_t spanStart, spanStop;
  int64_t corrLen;
  float   temporalWeight;
  int32_t numBreaks;
  int64_t minSpacing;
} sgz_segm_config;

/* SelfSimilarity.ConfigLike, Api/SelfSimilarity.scala:61-143 */
typedef struct {
  int32_t stepSize;
  int32_t hasStart, hasStop;
  int64_t spanStart, spanStop;
  int64_t corrLen;
  int32_t decimation;
  float   temporalWeight;
  int32_t colorInv;
  float   colorWarp, colorCeil;
  const int32_t *lut;                   /* NULL = GrayScale; else PsychoOptical-style LUT  */
  int32_t lutSize;                      /* (IntensityPalette is third-party: host passes it) */
  int32_t precise;                      /* 0 = tiled FP32 Gram (sims within 1e-5 of the reference, grey within
                                           1 LSB); 1 = per-cell FP64 replay of the reference's arithmetic
                                           (bit-identical sims, pixel-identical image, ~30x slower) */
} sgz_self_config;

typedef struct {
  int32_t imgExt;       /* image is imgExt x imgExt */
  int32_t decim;        /* effective decimation after the 0xB504 auto-adjust */
  int32_t numCorrs;
  int32_t afStart;
  int64_t numCells;     /* unique cells computed = imgExt*(imgExt+1)/2 */
} sgz_self_geometry;

/* ------------------------------------------------------------------------------------------
 * library / context
 * ---------------------------------------------------------------------------------------- */
int         sgz_abi_version(void);
const char *sgz_last_error(void);                    /* thread-local, never NULL */
int         sgz_device_count(int32_t *count);        /* number of usable sm_100 devices */

int sgz_ctx_create(int32_t device, sgz_ctx **out);
int sgz_ctx_destroy(sgz_ctx *ctx);
int sgz_ctx_synchronize(sgz_ctx *ctx);
/* raw cudaStream_t of the context, so that a host can bracket calls with its own CUDA events */
void *sgz_ctx_stream(sgz_ctx *ctx);
/* device time in ms of the kernels launched by the most recent compute call on this context
 * (CUDA events on the context's stream), and how many kernels that call launched */
int sgz_ctx_last_timing(sgz_ctx *ctx, float *ms, int64_t *launches);
/* total kernels launched on this context so far */
int64_t sgz_ctx_launch_count(sgz_ctx *ctx);
/* Device buffers released by destroy calls are parked in a per-device pool and reused by the next
 * database / job of a similar size (a search re-created per query does not pay cudaMalloc/cudaFree of
 * multi-GB buffers again).  sgz_ctx_trim hands the parked memory of this context's device back to the
 * driver; freedBytes may be NULL.  Pool size limit: environment SGZ_POOL_MAX_GB (default 64). */
int sgz_ctx_trim(sgz_ctx *ctx, int64_t *freedBytes);

/* ------------------------------------------------------------------------------------------
 * feature database  (replaces: DB discovery + per-offset AudioFile.read + MathUtil.normalize,
 * FeatureCorrelationImpl.scala:42-71,169,195-197)
 * ---------------------------------------------------------------------------------------- */
/* numCh = numCoeffs + 1 (channel 0 = loudness).  norm = feat_norms.aif content as
 * [numCh][2] = {min, max} (FeatureCorrelationImpl.scala:61-71) or NULL for normalize = false. */
int sgz_db_create(sgz_ctx *ctx, int32_t numCh, const float *norm, sgz_db **out);
int sgz_db_destroy(sgz_db *db);
/* optional capacity hint: avoids re-allocation while files are added */
int sgz_db_reserve(sgz_db *db, int64_t totalFrames, int32_t numFiles);
/* Adds one feature file; the host buffer may be freed on return.  Files are searched in add
 * order (the reference iterates a HashSet; the host decides the order).  Returns the file
 * index >= 0, or an error code < 0. */
int sgz_db_add_file(sgz_db *db, const void *frames, int64_t nFrames, int32_t layout);
/* Same, but the frames already live in device memory of this context (interleaved LE). */
int sgz_db_add_file_device(sgz_db *db, const void *dFrames, int64_t nFrames);
/* Synthetic feature file generated on the device (bench / tests):
 * x[c][t] = mu[c] + sigma[c] * g(seed, stream, c, t), channel 0 clamped to >= floor0. */
int sgz_db_add_synth(sgz_db *db, uint64_t seed, uint32_t stream, int64_t nFrames,
                     const float *mu, const float *sigma, float floor0);
/* `numFiles` synthetic files of `nFramesEach` frames (streams firstStream, firstStream + 1, ...) generated by ONE
 * kernel launch; identical to numFiles calls of sgz_db_add_synth.  Returns the index of the first file added. */
int sgz_db_add_synth_many(sgz_db *db, uint64_t seed, uint32_t firstStream, int32_t numFiles, int64_t nFramesEach,
                          const float *mu, const float *sigma, float floor0);
/* Overwrites frames [frameOff, frameOff+n) of an added file with raw (un-normalised)
 * interleaved LE frames -- used to plant known needles into a synthetic DB. */
int sgz_db_patch(sgz_db *db, int32_t file, int64_t frameOff, const float *frames, int64_t n);
/* Ends the add phase: everything is normalised, planar and resident in HBM afterwards. */
int sgz_db_finalize(sgz_db *db);
/* Ends the add phase WITHOUT waiting for uploads still in flight (SGZ_LAYOUT_HOST_STABLE files).  The first
 * search on the database then streams: the correlation kernel runs range by range behind the upload, like the
 * reference's loop that reads one file and correlates it (FeatureCorrelationImpl.scala:161-199), so the scan
 * hides behind the PCIe transfer.  Calling sgz_db_finalize afterwards is the explicit wait. */
int sgz_db_finalize_async(sgz_db *db);
/* FeatureStats (replaces FeatureStatsImpl.body(), Impl/FeatureStatsImpl.scala:30-135; the producer of
 * feat_norms.aif, Strugatzki.scala:411-430).  The database must hold RAW features (norm = NULL at create).
 * out[numCh][2] = per channel (min over files of the file's 1st percentile, max over files of its 99th) as
 * Doubles -- the host narrows them to Float when it writes the norm file; perFile (optional) =
 * [numFiles][numCh][2]. */
int sgz_db_stats(sgz_db *db, double *out, double *perFile);
int sgz_db_info(sgz_db *db, int32_t *numFiles, int64_t *totalFrames, int32_t *numCh);
int sgz_db_file_frames(sgz_db *db, int32_t file, int64_t *nFrames);
/* Reads back NORMALISED frames [frameOff, frameOff+n) of a file as planar [numCh][n]. */
int sgz_db_read(sgz_db *db, int32_t file, int64_t frameOff, int64_t n, float *out);

/* ------------------------------------------------------------------------------------------
 * FeatureCorrelation  (replaces FeatureCorrelationImpl.body(), :32-412)
 * ---------------------------------------------------------------------------------------- */
/* input = raw feature frames of config.metaInput's feature file (layout as above); the
 * punch-in / punch-out windows are cut, normalised and analysed exactly like readInBuffer
 * (FeatureCorrelationImpl.scala:83-98). */
int sgz_corr_create(sgz_db *db, const sgz_corr_config *cfg, const void *input,
                    int64_t inputFrames, int32_t layout, sgz_corr **out);
int sgz_corr_destroy(sgz_corr *job);

/* One-call synchronous search on this context's shard: scan + select + merge.
 * Punch-in searches (no punch-out): the matches are the reference's with the reference's sims BIT FOR BIT -- every
 * offset that can reach the result is re-evaluated in the reference's Double arithmetic (MathUtil.stat / correlate,
 * :29-62,177-196) before the queues of addMatch (:120-150) are replayed, so equal sims of repeated material are one
 * entry like in the TreeSet and the result does not depend on how the database is sharded.  Punch-out searches:
 * same files and spans, sims within 1e-5 relative. */
int sgz_corr_run(sgz_corr *job);
/* Asynchronous variant: runs on a worker thread; poll/abort from the host loop
 * (reference: checkAborted() / progress_=, FeatureCorrelationImpl.scala:164,192,402). */
int sgz_corr_start(sgz_corr *job);
int sgz_corr_poll(sgz_corr *job, float *progress, int32_t *done, int32_t *status);
int sgz_corr_abort(sgz_corr *job);
int sgz_corr_wait(sgz_corr *job);

/* Results (descending sim, Float.compare order, like allPrio.toIndexedSeq :410). */
int sgz_corr_result(sgz_corr *job, sgz_match *out, int32_t cap, int32_t *n);
/* number of punch-in frame-offsets evaluated by the scan (the benchmark unit) */
int sgz_corr_num_offsets(sgz_corr *job, int64_t *n);

/* device time (CUDA events on the context stream) of the last scan (K1) and of the selection
 * kernels (K2) accumulated since that scan, plus the number of kernels the scan launched.
 * Punch-in searches on the tensor-core scan: scanMs is the K1 launch alone; what follows it inside
 * sgz_corr_scan (exact re-evaluation of ill-conditioned and of decisive offsets, per-file boosts)
 * is the first part of selectMs. */
int sgz_corr_timing(sgz_corr *job, float *scanMs, float *selectMs, int64_t *scanLaunches);

/* Debug / parity access to the ungated curves in HBM: which = 0 punch-in, 1 punch-out.
 * Writes sim/boost of window starts [first, first+n) of `file` (local frame indices). */
int sgz_corr_curve(sgz_corr *job, int32_t which, int32_t file, int64_t first, int64_t n,
                   float *sim, float *boost);

/* ---- phase-wise protocol for a database sharded over several contexts / processes --------
 * Every rank holds a contiguous range of the global (ordered) file list.  The host moves the
 * small exchange buffers between ranks (torch.distributed / NCCL all_gather); the library
 * never talks to another process itself.  Protocol (identical on every rank):
 *
 *   sgz_corr_scan(job)                          K1 on the local shard
 *   sgz_corr_local_summary(job, buf)            per-file maxima of the local shard
 *   -- all_gather summaries, concatenated in rank order --
 *   sgz_corr_set_global(job, all, nFilesGlobal, myFirstFile)
 *   loop:
 *     sgz_corr_select(job, &nRec)               local candidate records for this round
 *     sgz_corr_records(job, buf, cap)           (fixed-size POD records)
 *     -- all_gather records, concatenated in rank order --
 *     sgz_corr_merge(job, allRecs, nAll, &done) replicated, deterministic replay of the
 *                                               reference's addMatch state machine
 *   until done
 *   sgz_corr_result(job, ...)                   identical on every rank; Match.file is GLOBAL
 */
typedef struct {
  float   maxSim;       /* largest non-NaN sim of the file's punch-in curve (-inf if none)   */
  int32_t numOffsets;   /* evaluated punch-in offsets of the file                            */
  float   maxSimOut;    /* same for the punch-out curve (punch-out mode only, else -inf)     */
  int32_t _pad;
} sgz_file_summary;

typedef struct {
  int32_t file;         /* GLOBAL file index                                   */
  int32_t kind;         /* 0 = candidate cell, 1 = resolved entry of a filling-phase file,
                           2 = end-of-file marker of a resolved file,
                           3 = punch-in offset above the round threshold (punch-out mode),
                           4 = gate interval of a filling-round file in punch-out mode: piOff = first
                               accepted row (-1: none), sim = its in-sim, boostIn = largest in-sim the row
                               gate turned away before it; sgz_corr_merge ends the round at the first
                               file whose interval does not hold allPrio.last.sim^2 of that moment       */
  int32_t piOff;        /* punch-in offset (feature frames, file local)        */
  int32_t poOff;        /* punch-out offset, or -1                              */
  float   sim;
  float   boostIn;
  float   boostOut;
  int32_t aux;
} sgz_record;

/* Punch-in-only searches need far less than one summary per file: a file whose maximum is not among the numMatches largest
 * maxima of its rank can never raise the threshold of a later file, and whether a file is scanned for candidates is
 * decided by the rank that holds it.  sgz_corr_local_top returns those (at most numMatches) entries with LOCAL file
 * indices and the number of local files; after an all_gather the host rebases the indices to the global file list and
 * hands all entries of all ranks to sgz_corr_set_global_top (in place of local_summary / set_global: 8 B x numMatches
 * per rank instead of 16 B x files).  Thresholds derived from a subset of the earlier files are lower bounds of the
 * exact ones, i.e. safe pre-filters (select.cuh); the merge sees the same records on every rank. */
typedef struct {
  int32_t file;
  float   maxSim;
} sgz_file_entry;
int sgz_corr_local_top(sgz_corr *job, sgz_file_entry *out, int32_t cap, int32_t *n, int32_t *numFiles);

/* numPerFile = 1, punch-in only: the entry of a file is its maximum at its first position (addMatch keeps one match per
 * file and replaces it only by a strictly larger sim, FeatureCorrelationImpl.scala:135-150), and allPrio always holds the
 * numMatches largest entries seen so far (:120-129, :399-400) -- so the entries of each rank's numMatches best files are ALL a
 * search needs, and ONE exchange per search replaces the summary / record rounds above.
 * sgz_corr_local_best returns those entries as records (kind 1, LOCAL file index, punch-in offset, sim, boostIn) and says
 * whether the shortcut applies on this rank: *ok = 0 when the scan did not prepare it (another kernel, numPerFile > 1, a
 * punch-out search) or when a file of this rank holds a NaN window (whose entry the reference decides by its NaN-first
 * rule while allPrio still has space).  If every rank says ok, the host rebases the file indices, hands all records of
 * all ranks to sgz_corr_finish_from_best on every rank -- the search is then finished, sgz_corr_result is valid -- else it
 * continues with sgz_corr_local_top as before. */
int sgz_corr_local_best(sgz_corr *job, sgz_record *out, int32_t cap, int32_t *n, int32_t *numFiles, int32_t *ok);
int sgz_corr_finish_from_best(sgz_corr *job, const sgz_record *all, int32_t nAll, int32_t nFilesGlobal);
int sgz_corr_set_global_top(sgz_corr *job, const sgz_file_entry *all, int32_t nAll, int32_t nFilesGlobal,
                            int32_t myFirstFile);

int sgz_corr_scan(sgz_corr *job);
int sgz_corr_local_summary(sgz_corr *job, sgz_file_summary *out, int32_t cap, int32_t *n);
int sgz_corr_set_global(sgz_corr *job, const sgz_file_summary *all, int32_t nFilesGlobal,
                        int32_t myFirstFile);
int sgz_corr_select(sgz_corr *job, int32_t *nRecords);
int sgz_corr_records(sgz_corr *job, sgz_record *out, int32_t cap, int32_t *n);
int sgz_corr_merge(sgz_corr *job, const sgz_record *all, int32_t nAll, int32_t *done);

/* ------------------------------------------------------------------------------------------
 * FeatureSegmentation  (replaces FeatureSegmentationImpl.body(), :31-142)
 * ---------------------------------------------------------------------------------------- */
/* frames = raw feature frames of metaInput's feature file.  curve (optional, may be NULL)
 * receives the sim of every offset (afLen - 2H + 1 values). */
int sgz_segm_run(sgz_ctx *ctx, const sgz_segm_config *cfg, int32_t numCh, const float *norm,
                 const void *frames, int64_t nFrames, int32_t layout, sgz_break *out,
                 int32_t cap, int32_t *n, float *curve, int64_t curveCap, int64_t *numOffsets);

/* ------------------------------------------------------------------------------------------
 * SelfSimilarity  (replaces SelfSimilarityImpl.body(), :31-180, up to the ImageIO.write)
 * ---------------------------------------------------------------------------------------- */
int sgz_self_geometry_of(const sgz_self_config *cfg, int64_t nFrames1, int64_t nFrames2,
                         sgz_self_geometry *out);
/* frames2 = NULL for plain self similarity (metaInput2 = None).  rgb (may be NULL to skip the
 * download) receives imgExt*imgExt packed 0x00RRGGBB pixels, row-major, like
 * BufferedImage.TYPE_INT_RGB (SelfSimilarityImpl.scala:117-118,152-155).
 * rowBegin/rowEnd select the block of image COLUMNS x in [rowBegin,rowEnd) (= leftOff/decim)
 * this call computes (multi-GPU sharding by row-tile blocks); 0,imgExt = everything. */
int sgz_self_run(sgz_ctx *ctx, const sgz_self_config *cfg, int32_t numCh, const float *norm,
                 const void *frames1, int64_t nFrames1, const void *frames2, int64_t nFrames2,
                 int32_t layout, int32_t rowBegin, int32_t rowEnd, int32_t *rgb, int64_t rgbCap,
                 sgz_self_geometry *geom);
/* which kernel rendered the last image of this context: 0 none, 1 FP32 FFMA2 Gram tiles (selfsim_fast.cuh),
 * 2 tensor-core Gram tiles (selfsim_tc.cuh: split-FP16 tcgen05, the default when the geometry fits), 3 FP64 replay */
int32_t sgz_self_last_kernel(sgz_ctx *ctx);
/* sims of selected cells (decimated image coordinates), for parity checks; with precise = 0 and imgExt <= 4096 the
 * cells of the computed triangle are read back from the image kernel itself */
int sgz_self_cells(sgz_ctx *ctx, const sgz_self_config *cfg, int32_t numCh, const float *norm,
                   const void *frames1, int64_t nFrames1, const void *frames2, int64_t nFrames2,
                   int32_t layout, int64_t nCells, const int32_t *leftIdx,
                   const int32_t *rightIdx, float *sim, int32_t *rgb);

/* ------------------------------------------------------------------------------------------
 * CrossSimilarity  (replaces CrossSimilarityImpl.body(), Impl/CrossSimilarityImpl.scala:32-187)
 * The shorter span is the template, slid over the longer one; one sim per step -- the values the reference
 * writes into its 1-channel output audio file.  The reference's ring-buffer behaviour (8192-frame buffer whose
 * first read takes min(len, 8192) frames, write index wrapping at the template length, read index wrapping
 * at 8192) is reproduced as it is: 1 + len2 - min(len2, 8192) values, bit-identical Double arithmetic.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t stepSize;
  int32_t has1Start, has1Stop;   /* span1: Span.All when both are 0 */
  int32_t has2Start, has2Stop;   /* span2 */
  int32_t _pad;
  int64_t span1Start, span1Stop; /* sample frames */
  int64_t span2Start, span2Stop;
  float   temporalWeight;
  float   maxBoost;
} sgz_cross_config;

/* number of values a run writes; SGZ_ERR_INVALID where the reference throws (shorter span > 8192 feature
 * frames: ArrayIndexOutOfBounds; shorter span empty: ArithmeticException) */
int sgz_cross_num_outputs(const sgz_cross_config *cfg, int64_t nFrames1, int64_t nFrames2, int64_t *nOut);
/* frames1 / frames2: raw feature frames of metaInput1 / metaInput2 in `layout`; sim[simCap] may be NULL */
int sgz_cross_run(sgz_ctx *ctx, const sgz_cross_config *cfg, int32_t numCh, const float *norm,
                  const void *frames1, int64_t nFrames1, const void *frames2, int64_t nFrames2,
                  int32_t layout, float *sim, int64_t simCap, int64_t *nOut);

/* ------------------------------------------------------------------------------------------
 * measurement helpers (bench.py): live peaks of the pipes the kernels are bound by
 * ---------------------------------------------------------------------------------------- */
/* which: 0 = FP32 FFMA TFLOP/s, 1 = packed FFMA2 TFLOP/s, 2 = FP64 DFMA TFLOP/s,
 *        3 = HBM copy GB/s (read+write), 4 = shared-memory LDS.128 GB/s */
int sgz_measure_peak(sgz_ctx *ctx, int32_t which, double *value);

#ifdef __cplusplus
}
#endif
#endif
