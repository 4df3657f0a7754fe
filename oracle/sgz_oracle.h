/*
 * sgz_oracle.h -- CPU ORACLE for the Strugatzki feature-similarity hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product
 * (strugatzki_b200/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C, single-threaded, Double-accumulating restatement of the reference's
 * Scala loops, following their loop order, ring-buffer indexing and float/double
 * conversion points:
 *   Impl/MathUtil.scala:29-62 (stat), :80-99 (correlateHalf), :109-118 (avg),
 *   :132-152 (normalize), :177-196 (correlate)
 *   Impl/FeatureCorrelationImpl.scala:32-421
 *   Impl/FeatureSegmentationImpl.scala:31-142
 *   Impl/SelfSimilarityImpl.scala:31-180
 *   Impl/CrossSimilarityImpl.scala:32-187
 *   Impl/FeatureStatsImpl.scala:30-135
 *   Impl/SpanUtil.scala:38-43, Api/FeatureCorrelation.scala:75-77,
 *   Api/FeatureSegmentation.scala:60-62
 *
 * PARITY UNPINNED: the reference ships no numeric test, golden vector or fixture for this
 * path (src/test/.../StrugatzkiSuite.scala only round-trips Config XML) and no JVM exists in
 * the build container, so this restatement cannot be checked against the reference's own
 * output.  It is cross-checked against an independent numpy closed-form formulation
 * (tests/test_oracle.py) and against committed fixtures generated from itself
 * (tests/golden/).
 *
 * Documented deviations from the JVM reference:
 *   Q5  punch-out tail: cells that the reference reads past the written part of its temp
 *       files (stale data of earlier DB files) are defined as "not searched".
 *   Q6  DB files shorter than the punch window produce no offsets (the reference evaluates
 *       one offset on a ring buffer that still holds the previous file's frames).
 *   DB iteration order is the order of the file list handed in (the reference iterates a
 *   HashSet of java.io.File).
 */
#ifndef SGZ_ORACLE_H
#define SGZ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  float   sim;
  int32_t file;      /* index into the DB file list */
  int64_t start;     /* sample frames (feature index * stepSize) */
  int64_t stop;
  float   boostIn;
  float   boostOut;
} sgz_o_match;

typedef struct {
  float   sim;
  int32_t _pad;
  int64_t pos;       /* sample frames */
} sgz_o_break;

typedef struct {
  int32_t numCh;          /* numCoeffs + 1; channel 0 = loudness */
  int32_t stepSize;       /* fftSize / fftOverlap */
  const float *norm;      /* [numCh][2] = {min,max} or NULL (normalize = false) */
  const float *input;     /* feature frames of metaInput, interleaved [frame][numCh] */
  int64_t inputFrames;
  int64_t punchInStart, punchInStop;     /* sample frames */
  float   punchInWeight;
  int32_t hasPunchOut;
  int64_t punchOutStart, punchOutStop;
  float   punchOutWeight;
  int64_t minPunch, maxPunch;            /* sample frames */
  float   maxBoost;
  int32_t numMatches, numPerFile;
  int64_t minSpacing;                    /* sample frames */
} sgz_o_corr_cfg;

typedef struct {
  int32_t numCh;
  int32_t stepSize;
  const float *norm;
  int32_t hasStart, hasStop;             /* Span.HasStart / HasStop */
  int64_t spanStart, spanStop;           /* sample frames */
  int64_t corrLen;                       /* sample frames; HALF window */
  float   temporalWeight;
  int32_t numBreaks;
  int64_t minSpacing;
} sgz_o_segm_cfg;

typedef struct {
  int32_t numCh;
  int32_t stepSize;
  const float *norm;
  int32_t hasStart, hasStop;
  int64_t spanStart, spanStop;
  int64_t corrLen;
  int32_t decimation;
  float   temporalWeight;
  int32_t colorInv;                      /* GrayScale only (IntensityPalette LUT not available) */
  float   colorWarp, colorCeil;
} sgz_o_self_cfg;

/* ---- MathUtil ---- */
void  sgz_o_stat(const float *const *mat, int frameOff, int frameLen, int chanOff, int chanLen,
                 double *mean, double *stddev);
float sgz_o_correlate_half(int numChannels, int halfWinSize, const float *const *a, int frameOff,
                           int chanOff);
float sgz_o_avg(const float *b, int off, int len);
void  sgz_o_normalize(const float *norm, float *const *b, int numCh, int bOff, int bLen);
float sgz_o_correlate(const float *const *a, double aMean, double aStdDev, int numFrames,
                      int numChannels, const float *const *b, int bLen, double bMean,
                      double bStdDev, int bFrameOff, int bChanOff);

/* ---- FeatureCorrelationImpl.body() ---- */
/* Full search.  files[i] = interleaved raw frames of DB file i.  Returns number of matches
 * written (descending sim, Float.compare order), or <0 on error. */
int sgz_o_corr_search(const sgz_o_corr_cfg *cfg, int numFiles, const float *const *files,
                      const int64_t *nFrames, sgz_o_match *out, int cap);

/* Loop A / loop B curve of ONE file without any gating: sim and boost for every window start
 * in [firstFrame, nFrames - W].  which = 0 punch-in window, 1 punch-out window.  The ring
 * buffer is primed at firstFrame exactly like the reference primes it at 0 (loop A) or
 * poOff0 (loop B).  Returns the number of values written. */
int64_t sgz_o_corr_curve(const sgz_o_corr_cfg *cfg, int which, const float *file,
                         int64_t nFrames, int64_t firstFrame, float *sim, float *boost,
                         int64_t cap);

/* number of evaluated punch-in offsets of the whole DB (metric unit) */
int64_t sgz_o_corr_num_offsets(const sgz_o_corr_cfg *cfg, int numFiles, const int64_t *nFrames);

/* ---- FeatureSegmentationImpl.body() ---- */
int sgz_o_segm_run(const sgz_o_segm_cfg *cfg, const float *file, int64_t nFrames,
                   sgz_o_break *out, int cap, float *curve /* optional, afLen-2H+1 */,
                   int64_t curveCap);

/* ---- SelfSimilarityImpl.body() ---- */
/* Geometry only: returns imgExt, fills decim / numCorrs. */
int sgz_o_self_geometry(const sgz_o_self_cfg *cfg, int64_t nFrames1, int64_t nFrames2,
                        int32_t *decim, int32_t *numCorrs, int32_t *afStart);
/* Full image (imgExt*imgExt packed 0x00RRGGBB), row-major like BufferedImage.TYPE_INT_RGB. */
int sgz_o_self_image(const sgz_o_self_cfg *cfg, const float *file1, int64_t nFrames1,
                     const float *file2, int64_t nFrames2, int32_t *rgb, int64_t cap);
/* Selected cells: (leftIdx[i], rightIdx[i]) in decimated image coordinates -> sim and colour */
int sgz_o_self_cells(const sgz_o_self_cfg *cfg, const float *file1, int64_t nFrames1,
                     const float *file2, int64_t nFrames2, int64_t nCells,
                     const int32_t *leftIdx, const int32_t *rightIdx, float *sim, int32_t *rgb);

/* ---- CrossSimilarityImpl.body(), Impl/CrossSimilarityImpl.scala:32-187 ----
 * The reference is restated AS IT BEHAVES, including its ring-buffer quirks:
 *  - the buffer has 8192 frames; the FIRST read takes min(len2, 8192) frames, every later read one
 *    frame, written at readOff which wraps modulo len1 (:140-141,165);
 *  - MathUtil.correlate wraps its read index modulo the BUFFER length 8192 (MathUtil.scala:189), not
 *    modulo len1, so logical frame i of output k is buffer position (i + k % len1) % 8192;
 *  - mean / std-dev / loudness average of "b" always come from buffer positions [0, len1) (:181-185);
 *  - hence 1 + len2 - min(len2, 8192) output values.
 * len1 > 8192 fails in the reference with ArrayIndexOutOfBounds -> returns -2 here. */
typedef struct {
  int32_t numCh;
  int32_t stepSize;
  const float *norm;
  int32_t has1Start, has1Stop;           /* span1 */
  int64_t span1Start, span1Stop;
  int32_t has2Start, has2Stop;           /* span2 */
  int64_t span2Start, span2Stop;
  float   temporalWeight;
  float   maxBoost;
} sgz_o_cross_cfg;

/* returns the number of output values (written up to cap), <0 on error */
int64_t sgz_o_cross_run(const sgz_o_cross_cfg *cfg, const float *file1, int64_t nFrames1,
                        const float *file2, int64_t nFrames2, float *sim, int64_t cap);

/* ---- FeatureStatsImpl.body(), Impl/FeatureStatsImpl.scala:30-135 ----
 * Per channel: (min over files of the file's 1st percentile, max over files of its 99th), the percentiles taken
 * from a 2048-bin histogram of the skew-warped values.  files[i] = interleaved RAW frames.  out = [numCh][2]
 * doubles (the reference narrows them to Float when it writes feat_norms.aif, Strugatzki.scala:417-426);
 * perFile (optional) = [numFiles][numCh][2].  math.pow / math.log are libm's here (Java's are within 1 ulp of
 * them, not bit-identical: parity unpinned like the rest). */
int sgz_o_stats_run(int numCh, int numFiles, const float *const *files, const int64_t *nFrames, double *out,
                    double *perFile);

#ifdef __cplusplus
}
#endif
#endif
