/*
 * sgz_oracle.c -- CPU ORACLE (test infrastructure; see sgz_oracle.h for the contract).
 *
 * Plain C restatement of the reference's hot path.  Every function cites the reference
 * file:line it follows.  Build with -ffp-contract=off so that no float/double operation is
 * fused (the JVM never fuses).  PARITY UNPINNED (no reference fixtures exist; see header).
 */
#include "sgz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Java semantics helpers
 * ---------------------------------------------------------------------------------------- */

/* java.lang.Float.floatToIntBits: canonical NaN */
static int32_t j_float_bits(float x) {
  int32_t b;
  if (x != x) return 0x7fc00000;
  memcpy(&b, &x, 4);
  return b;
}

/* java.lang.Float.compare: total order, -0.0 < +0.0, NaN greatest and equal to itself */
static int j_float_compare(float x, float y) {
  if (x < y) return -1;
  if (x > y) return 1;
  int32_t a = j_float_bits(x), b = j_float_bits(y);
  return (a == b) ? 0 : (a < b ? -1 : 1);
}

/* Scala Double.toInt: truncate, saturate, NaN -> 0 */
static int32_t j_d2i(double d) {
  if (d != d) return 0;
  if (d >= 2147483647.0) return 2147483647;
  if (d <= -2147483648.0) return (-2147483647 - 1);
  return (int32_t)d;
}

/* ------------------------------------------------------------------------------------------
 * MathUtil  (Impl/MathUtil.scala)
 * ---------------------------------------------------------------------------------------- */

/* MathUtil.stat, Impl/MathUtil.scala:29-62 -- matrix-wide mean / population stddev, 2 passes */
void sgz_o_stat(const float *const *mat, int frameOff, int frameLen, int chanOff, int chanLen,
                double *mean_out, double *stddev_out) {
  int chanStop = chanOff + chanLen, frameStop = frameOff + frameLen;
  double sum = 0.0;
  for (int ch = chanOff; ch < chanStop; ch++) {
    const float *cb = mat[ch];
    for (int i = frameOff; i < frameStop; i++) sum += cb[i];
  }
  int matSize = frameLen * chanLen;
  double mean = sum / matSize;
  sum = 0.0;
  for (int ch = chanOff; ch < chanStop; ch++) {
    const float *cb = mat[ch];
    for (int i = frameOff; i < frameStop; i++) {
      double d = cb[i] - mean;
      sum += d * d;
    }
  }
  *mean_out = mean;
  *stddev_out = sqrt(sum / matSize);
}

/* MathUtil.correlateHalf, Impl/MathUtil.scala:80-99 */
float sgz_o_correlate_half(int numChannels, int halfWinSize, const float *const *a, int frameOff,
                           int chanOff) {
  int numFrames = halfWinSize << 1;
  double mean, stdDev;
  sgz_o_stat(a, 0, numFrames, chanOff, numChannels, &mean, &stdDev);
  double add = -mean;
  int matSize = numChannels * halfWinSize;
  double sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const float *ca = a[ch + chanOff];
    int j = frameOff + halfWinSize;
    for (int i = frameOff; i < j; i++) {
      sum += (ca[i % numFrames] + add) * (ca[(i + halfWinSize) % numFrames] + add);
    }
  }
  return (float)(sum / (stdDev * stdDev * matSize));
}

/* MathUtil.avg, Impl/MathUtil.scala:109-118 */
float sgz_o_avg(const float *b, int off, int len) {
  double sum = 0.0;
  for (int i = off; i < off + len; i++) sum += b[i];
  return (float)(sum / len);
}

/* MathUtil.normalize, Impl/MathUtil.scala:132-152; norm = [numCh][2] or NULL */
void sgz_o_normalize(const float *norm, float *const *b, int numCh, int bOff, int bLen) {
  if (norm == NULL) return;
  for (int ch = 0; ch < numCh; ch++) {
    float *cb = b[ch];
    float mn = norm[2 * ch], mx = norm[2 * ch + 1];
    float d = mx - mn;
    for (int i = bOff; i < bOff + bLen; i++) {
      float f = cb[i];
      cb[i] = (f - mn) / d;
    }
  }
}

/* MathUtil.correlate, Impl/MathUtil.scala:177-196; bLen = cb.length (ring length) */
float sgz_o_correlate(const float *const *a, double aMean, double aStdDev, int numFrames,
                      int numChannels, const float *const *b, int bLen, double bMean,
                      double bStdDev, int bFrameOff, int bChanOff) {
  double aAdd = -aMean, bAdd = -bMean;
  int aMatSize = numChannels * numFrames;
  double sum = 0.0;
  for (int ch = 0; ch < numChannels; ch++) {
    const float *ca = a[ch];
    const float *cb = b[ch + bChanOff];
    for (int i = 0; i < numFrames; i++) {
      sum += (ca[i] + aAdd) * (cb[(i + bFrameOff) % bLen] + bAdd);
    }
  }
  return (float)(sum / (aStdDev * bStdDev * aMatSize));
}

/* ------------------------------------------------------------------------------------------
 * FeatureCorrelation: input matrix, sorted sets, body()
 * ---------------------------------------------------------------------------------------- */

typedef struct {           /* FeatureMatrix, Api/FeatureCorrelation.scala:279-283 */
  float **mat;             /* [numChannels][numFrames] */
  int numFrames, numChannels;
  double mean, stdDev;
} feat_matrix;

typedef struct {           /* InputMatrix, Api/FeatureCorrelation.scala:285-289 */
  float **all;             /* [numCh][numFrames] backing store */
  feat_matrix temporal, spectral;
  double lnAvgLoudness;
  int numFrames;
} input_matrix;

static float **alloc_planar(int numCh, int len) {
  float **p = (float **)calloc((size_t)numCh, sizeof(float *));
  for (int c = 0; c < numCh; c++) p[c] = (float *)calloc((size_t)(len > 0 ? len : 1), sizeof(float));
  return p;
}
static void free_planar(float **p, int numCh) {
  if (!p) return;
  for (int c = 0; c < numCh; c++) free(p[c]);
  free(p);
}

/* AudioFile.read(buf, off, len) of interleaved source frames into a planar buffer */
static void read_frames(const float *file, int numCh, int64_t pos, float *const *buf, int off,
                        int len) {
  for (int i = 0; i < len; i++) {
    const float *fr = file + (pos + i) * numCh;
    for (int c = 0; c < numCh; c++) buf[c][off + i] = fr[c];
  }
}

static int full_to_feat(int64_t n, int step) { /* FeatureCorrelationImpl.scala:38 */
  return (int)((n + (step >> 1)) / step);
}
static int64_t feat_to_full(int i, int step) { /* :39 */
  return (int64_t)i * step;
}

/* calcLnAvgLoud, FeatureCorrelationImpl.scala:73 */
static double calc_ln_avg_loud(const float *b, int off, int len) { return log((double)sgz_o_avg(b, off, len)); }

/* readInBuffer, FeatureCorrelationImpl.scala:83-98.  returns 0 ok */
static int read_in_buffer(const sgz_o_corr_cfg *cfg, int64_t spanStart, int64_t spanStop,
                          input_matrix *m) {
  int step = cfg->stepSize, numCh = cfg->numCh;
  int start = full_to_feat(spanStart, step), stop = full_to_feat(spanStop, step);
  int frameNum = stop - start;
  if (frameNum <= 0 || start < 0 || stop > cfg->inputFrames) return -1;
  m->all = alloc_planar(numCh, frameNum);
  read_frames(cfg->input, numCh, start, m->all, 0, frameNum);
  sgz_o_normalize(cfg->norm, m->all, numCh, 0, frameNum);
  m->numFrames = frameNum;
  m->temporal.mat = m->all;          /* b.take(1) */
  m->temporal.numChannels = 1;
  m->temporal.numFrames = frameNum;
  sgz_o_stat((const float *const *)m->temporal.mat, 0, frameNum, 0, 1, &m->temporal.mean,
             &m->temporal.stdDev);
  m->spectral.mat = m->all + 1;      /* b.drop(1) */
  m->spectral.numChannels = numCh - 1;
  m->spectral.numFrames = frameNum;
  sgz_o_stat((const float *const *)m->spectral.mat, 0, frameNum, 0, numCh - 1, &m->spectral.mean,
             &m->spectral.stdDev);
  m->lnAvgLoudness = calc_ln_avg_loud(m->all[0], 0, frameNum);
  return 0;
}

/* private correlate, FeatureCorrelationImpl.scala:414-421 */
static float corr_wrap(const feat_matrix *a, const float *const *b, int bLen, int bFrameOff,
                       int bChanOff) {
  double bMean, bStdDev;
  sgz_o_stat(b, 0, a->numFrames, bChanOff, a->numChannels, &bMean, &bStdDev);
  return sgz_o_correlate((const float *const *)a->mat, a->mean, a->stdDev, a->numFrames,
                         a->numChannels, b, bLen, bMean, bStdDev, bFrameOff, bChanOff);
}

/* calcBoost, FeatureCorrelationImpl.scala:75-78 */
static float calc_boost(const input_matrix *in, const float *b) {
  double lnAvgB = calc_ln_avg_loud(b, 0, in->numFrames);
  return (float)exp((in->lnAvgLoudness - lnAvgB) / 0.6);
}

/* sim of one ring-buffer state, FeatureCorrelationImpl.scala:198-210 / :289-300 */
static float ring_sim(const input_matrix *in, float weight, float maxBoost, const float *const *ring,
                      int ringOff, float *boost_out) {
  float boost = calc_boost(in, ring[0]);
  float sim;
  if (boost <= maxBoost) {
    float temporal = (weight > 0.0f) ? corr_wrap(&in->temporal, ring, in->numFrames, ringOff, 0) : 0.0f;
    float spectral = (weight < 1.0f) ? corr_wrap(&in->spectral, ring, in->numFrames, ringOff, 1) : 0.0f;
    sim = temporal * weight + spectral * (1.0f - weight);
  } else {
    sim = 0.0f;
  }
  *boost_out = boost;
  return sim;
}

/* immutable.SortedSet[Match](MatchMinOrd): descending Float.compare order on sim only;
 * `+` never overwrites an element that compares equal; `-` removes the element comparing equal.
 * Api/FeatureCorrelation.scala:75-77 */
typedef struct {
  sgz_o_match *e;
  int n, cap;
} match_set;

static int mcmp(float a_sim, float b_sim) { return j_float_compare(b_sim, a_sim); } /* MatchMinOrd.compare(a,b) */

static void mset_init(match_set *s, int cap) {
  s->cap = cap < 8 ? 8 : cap;
  s->e = (sgz_o_match *)malloc(sizeof(sgz_o_match) * (size_t)s->cap);
  s->n = 0;
}
static void mset_free(match_set *s) { free(s->e); s->e = NULL; s->n = 0; }
static int mset_find(const match_set *s, float sim, int *found) {
  int i = 0;
  *found = 0;
  for (; i < s->n; i++) {
    int c = mcmp(sim, s->e[i].sim);   /* <0: sim sorts before e[i] */
    if (c == 0) { *found = 1; return i; }
    if (c < 0) return i;
  }
  return i;
}
static void mset_add(match_set *s, const sgz_o_match *m) {
  int found, i = mset_find(s, m->sim, &found);
  if (found) return;
  if (s->n == s->cap) {
    s->cap *= 2;
    s->e = (sgz_o_match *)realloc(s->e, sizeof(sgz_o_match) * (size_t)s->cap);
  }
  memmove(s->e + i + 1, s->e + i, sizeof(sgz_o_match) * (size_t)(s->n - i));
  s->e[i] = *m;
  s->n++;
}
static void mset_remove(match_set *s, float sim) {
  int found, i = mset_find(s, sim, &found);
  if (!found) return;
  memmove(s->e + i, s->e + i + 1, sizeof(sgz_o_match) * (size_t)(s->n - i - 1));
  s->n--;
}

/* SpanUtil.spacing, Impl/SpanUtil.scala:38-43 */
static int64_t span_spacing(int64_t aStart, int64_t aStop, int64_t bStart, int64_t bStop) {
  if (aStart < bStart) return bStart - aStop;
  return aStart - bStop;
}

typedef struct {
  const sgz_o_corr_cfg *cfg;
  match_set allPrio, entryPrio;
  int hasLast;
  sgz_o_match last;       /* lastEntryMatch */
} corr_state;

/* FeatureCorrelationImpl.scala:120-123 */
static int entry_has_space(const corr_state *st) {
  int a = st->cfg->numMatches - st->allPrio.n, b = st->cfg->numPerFile;
  int maxEntrySz = a < b ? a : b;
  return st->entryPrio.n < maxEntrySz;
}
/* :125-129 */
static float lowest_sim(const corr_state *st) {
  if (st->entryPrio.n > 0) return st->entryPrio.e[st->entryPrio.n - 1].sim;
  if (st->allPrio.n > 0) return st->allPrio.e[st->allPrio.n - 1].sim;
  return 0.0f;
}
/* :135-150 */
static void add_match(corr_state *st, const sgz_o_match *m) {
  if (st->hasLast &&
      span_spacing(m->start, m->stop, st->last.start, st->last.stop) < st->cfg->minSpacing) {
    if (st->last.sim < m->sim) {
      mset_remove(&st->entryPrio, st->last.sim);
      mset_add(&st->entryPrio, m);
      st->last = *m;
    }
  } else {
    mset_add(&st->entryPrio, m);
    if (st->entryPrio.n > st->cfg->numPerFile) {
      st->entryPrio.n--;           /* entryPrio -= entryPrio.last */
    }
    st->last = *m;
    st->hasLast = 1;
  }
}

static void free_input(input_matrix *m, int numCh) { free_planar(m->all, numCh); m->all = NULL; }

/* FeatureCorrelationImpl.body(), Impl/FeatureCorrelationImpl.scala:32-412 */
int sgz_o_corr_search(const sgz_o_corr_cfg *cfg, int numFiles, const float *const *files,
                      const int64_t *nFramesArr, sgz_o_match *out, int cap) {
  int step = cfg->stepSize, numCh = cfg->numCh;
  input_matrix matrixIn, matrixOut;
  memset(&matrixIn, 0, sizeof matrixIn);
  memset(&matrixOut, 0, sizeof matrixOut);
  if (read_in_buffer(cfg, cfg->punchInStart, cfg->punchInStop, &matrixIn) != 0) return -2;
  int hasOut = cfg->hasPunchOut != 0;
  if (hasOut && read_in_buffer(cfg, cfg->punchOutStart, cfg->punchOutStop, &matrixOut) != 0) {
    free_input(&matrixIn, numCh);
    return -2;
  }
  int punchInLen = matrixIn.numFrames;
  int punchOutLen = hasOut ? matrixOut.numFrames : 0;
  float inTempWeight = cfg->punchInWeight;

  corr_state st;
  st.cfg = cfg;
  mset_init(&st.allPrio, cfg->numMatches + cfg->numPerFile + 4);
  mset_init(&st.entryPrio, cfg->numPerFile + 4);
  st.hasLast = 0;

  int minPunch = full_to_feat(cfg->minPunch, step);
  int maxPunch = full_to_feat(cfg->maxPunch, step);

  float **eInBuf = alloc_planar(numCh, punchInLen);
  float **eOutBuf = alloc_planar(numCh, punchOutLen);
  float *tInSim = NULL, *tInBoost = NULL, *tOutSim = NULL, *tOutBoost = NULL;
  int64_t tCap = 0;

  for (int extrIdx = 0; extrIdx < numFiles; extrIdx++) {
    const float *file = files[extrIdx];
    int64_t numFrames = nFramesArr[extrIdx];

    st.entryPrio.n = 0;            /* :166 */
    st.hasLast = 0;                /* :167 */

    if (hasOut && numFrames + 1 > tCap) {
      tCap = numFrames + 1;
      tInSim = (float *)realloc(tInSim, sizeof(float) * (size_t)tCap);
      tInBoost = (float *)realloc(tInBoost, sizeof(float) * (size_t)tCap);
      tOutSim = (float *)realloc(tOutSim, sizeof(float) * (size_t)tCap);
      tOutBoost = (float *)realloc(tOutBoost, sizeof(float) * (size_t)tCap);
    }

    int tInOpen = 0;
    int tInOff = 0;
    int64_t tInCount = 0;
    int64_t left = numFrames;
    if (hasOut) left -= minPunch;  /* :183-185 */
    int readSz = punchInLen;
    int readOff = 0;
    int logicalOff = 0;
    int64_t filePos = 0;

    /* Q6 (documented deviation): a file that cannot fill the window yields no offsets */
    if (left < punchInLen) left = 0;

    while (left > 0) {             /* loop A, :190-246 */
      int chunkLen = (int)(left < readSz ? left : readSz);
      read_frames(file, numCh, filePos, eInBuf, readOff, chunkLen);
      filePos += chunkLen;
      int eInBufOff = logicalOff % punchInLen;
      sgz_o_normalize(cfg->norm, eInBuf, numCh, readOff, chunkLen);
      float boost;
      float sim = ring_sim(&matrixIn, inTempWeight, cfg->maxBoost, (const float *const *)eInBuf,
                           eInBufOff, &boost);

      if (hasOut) {
        if (tInOpen || entry_has_space(&st) || sim > lowest_sim(&st)) {
          if (!tInOpen) {
            tInOff = logicalOff;
            tInOpen = 1;
          }
          tInSim[tInCount] = sim;
          tInBoost[tInCount] = boost;
          tInCount++;
        }
      } else {
        if (entry_has_space(&st) || sim > lowest_sim(&st)) {
          sgz_o_match m;
          m.sim = sim;
          m.file = extrIdx;
          m.start = feat_to_full(logicalOff, step);
          m.stop = feat_to_full(logicalOff + punchInLen, step);
          m.boostIn = boost;
          m.boostOut = 1.0f;
          add_match(&st, &m);
        }
      }

      left -= chunkLen;
      readOff = (readOff + chunkLen) % punchInLen;
      logicalOff += 1;
      readSz = 1;
    }

    if (hasOut && tInOpen) {       /* :250-393 */
      int poOff0 = tInOff + minPunch;
      left = numFrames - poOff0;
      if (left >= punchOutLen) {
        float outTempWeight = cfg->punchOutWeight;
        filePos = poOff0;
        readSz = punchOutLen;
        readOff = 0;
        logicalOff = 0;
        int64_t tOutCount = 0;
        while (left > 0) {         /* loop B, :281-315 */
          int chunkLen = (int)(left < readSz ? left : readSz);
          read_frames(file, numCh, filePos, eOutBuf, readOff, chunkLen);
          filePos += chunkLen;
          sgz_o_normalize(cfg->norm, eOutBuf, numCh, readOff, chunkLen);
          int extraBufOff = logicalOff % punchOutLen;
          float boost;
          float sim = ring_sim(&matrixOut, outTempWeight, cfg->maxBoost,
                               (const float *const *)eOutBuf, extraBufOff, &boost);
          tOutSim[tOutCount] = sim;
          tOutBoost[tOutCount] = boost;
          tOutCount++;
          left -= chunkLen;
          readOff = (readOff + chunkLen) % punchOutLen;
          logicalOff += 1;
          readSz = 1;
        }

        /* loop C, :322-389.  Q5 (documented deviation): the reference iterates
         * numFrames - poOff0 times and bounds the inner loop by tOutSize = numFrames - poOff0,
         * reading past the tInCount / tOutCount values actually written; those cells are
         * "not searched" here. */
        int64_t nC = tInCount;
        int piOff = tInOff;
        for (int64_t k = 0; k < nC; k++) {
          float inSim = tInSim[k];
          float boostIn = tInBoost[k];
          float low = lowest_sim(&st);
          int hs = entry_has_space(&st);
          if (inSim > (low * low)) {   /* :342 */
            int poOff = piOff + minPunch;
            int64_t tOutSeek = piOff - tInOff;
            int64_t left2 = tOutCount - tOutSeek;
            int64_t span = (int64_t)maxPunch - minPunch + 1;
            if (span < left2) left2 = span;
            for (int64_t j = 0; j < left2; j++) {
              float outSim = tOutSim[tOutSeek + j];
              float boostOut = tOutBoost[tOutSeek + j];
              float prod = inSim * outSim;          /* Float * Float */
              float sim = (float)sqrt((double)prod); /* :370 */
              if (hs || sim > low) {
                sgz_o_match m;
                m.sim = sim;
                m.file = extrIdx;
                m.start = feat_to_full(piOff, step);
                m.stop = feat_to_full(poOff, step);
                m.boostIn = boostIn;
                m.boostOut = boostOut;
                add_match(&st, &m);
                low = lowest_sim(&st);
                hs = entry_has_space(&st);
              }
              poOff++;
            }
          }
          piOff++;
        }
      }
    }

    /* :399-400  allPrio ++= entryPrio; take(numMatches) */
    for (int i = 0; i < st.entryPrio.n; i++) mset_add(&st.allPrio, &st.entryPrio.e[i]);
    if (st.allPrio.n > cfg->numMatches) st.allPrio.n = cfg->numMatches < 0 ? 0 : cfg->numMatches;
  }

  int n = st.allPrio.n < cap ? st.allPrio.n : cap;
  for (int i = 0; i < n; i++) out[i] = st.allPrio.e[i];

  mset_free(&st.allPrio);
  mset_free(&st.entryPrio);
  free_planar(eInBuf, numCh);
  free_planar(eOutBuf, numCh);
  free(tInSim); free(tInBoost); free(tOutSim); free(tOutBoost);
  free_input(&matrixIn, numCh);
  if (hasOut) free_input(&matrixOut, numCh);
  return n;
}

/* ungated loop A / loop B curve of one file */
int64_t sgz_o_corr_curve(const sgz_o_corr_cfg *cfg, int which, const float *file, int64_t nFrames,
                         int64_t firstFrame, float *simOut, float *boostOut, int64_t cap) {
  int numCh = cfg->numCh;
  input_matrix in;
  memset(&in, 0, sizeof in);
  int rc = which == 0 ? read_in_buffer(cfg, cfg->punchInStart, cfg->punchInStop, &in)
                      : read_in_buffer(cfg, cfg->punchOutStart, cfg->punchOutStop, &in);
  if (rc != 0) return -2;
  float weight = which == 0 ? cfg->punchInWeight : cfg->punchOutWeight;
  int W = in.numFrames;
  float **ring = alloc_planar(numCh, W);
  int64_t left = nFrames - firstFrame;
  if (left < W) left = 0;
  int readSz = W, readOff = 0, logicalOff = 0;
  int64_t filePos = firstFrame, count = 0;
  while (left > 0 && count < cap) {
    int chunkLen = (int)(left < readSz ? left : readSz);
    read_frames(file, numCh, filePos, ring, readOff, chunkLen);
    filePos += chunkLen;
    sgz_o_normalize(cfg->norm, ring, numCh, readOff, chunkLen);
    float boost;
    float sim = ring_sim(&in, weight, cfg->maxBoost, (const float *const *)ring, logicalOff % W, &boost);
    simOut[count] = sim;
    boostOut[count] = boost;
    count++;
    left -= chunkLen;
    readOff = (readOff + chunkLen) % W;
    logicalOff++;
    readSz = 1;
  }
  free_planar(ring, numCh);
  free_input(&in, numCh);
  return count;
}

int64_t sgz_o_corr_num_offsets(const sgz_o_corr_cfg *cfg, int numFiles, const int64_t *nFrames) {
  int step = cfg->stepSize;
  int W = full_to_feat(cfg->punchInStop, step) - full_to_feat(cfg->punchInStart, step);
  int minPunch = cfg->hasPunchOut ? full_to_feat(cfg->minPunch, step) : 0;
  int64_t total = 0;
  for (int i = 0; i < numFiles; i++) {
    int64_t n = nFrames[i] - minPunch - W + 1;
    if (n > 0) total += n;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------
 * FeatureSegmentationImpl.body(), Impl/FeatureSegmentationImpl.scala:31-142
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  sgz_o_break *e;
  int n, cap;
} break_set;           /* SortedSet[Break](BreakMaxOrd): ascending Float.compare on sim */

static int bset_find(const break_set *s, float sim, int *found) {
  int i = 0;
  *found = 0;
  for (; i < s->n; i++) {
    int c = j_float_compare(sim, s->e[i].sim);
    if (c == 0) { *found = 1; return i; }
    if (c < 0) return i;
  }
  return i;
}
static void bset_add(break_set *s, const sgz_o_break *b) {
  int found, i = bset_find(s, b->sim, &found);
  if (found) return;
  if (s->n == s->cap) {
    s->cap *= 2;
    s->e = (sgz_o_break *)realloc(s->e, sizeof(sgz_o_break) * (size_t)s->cap);
  }
  memmove(s->e + i + 1, s->e + i, sizeof(sgz_o_break) * (size_t)(s->n - i));
  s->e[i] = *b;
  s->n++;
}
static void bset_remove(break_set *s, float sim) {
  int found, i = bset_find(s, sim, &found);
  if (!found) return;
  memmove(s->e + i, s->e + i + 1, sizeof(sgz_o_break) * (size_t)(s->n - i - 1));
  s->n--;
}

int sgz_o_segm_run(const sgz_o_segm_cfg *cfg, const float *file, int64_t nFrames, sgz_o_break *out,
                   int cap, float *curve, int64_t curveCap) {
  int step = cfg->stepSize, numCh = cfg->numCh, numCoeffs = cfg->numCh - 1;
  int halfWinLen = full_to_feat(cfg->corrLen, step);
  float tempWeight = cfg->temporalWeight;
  break_set prio;
  prio.cap = cfg->numBreaks + 8;
  prio.e = (sgz_o_break *)malloc(sizeof(sgz_o_break) * (size_t)prio.cap);
  prio.n = 0;
  int hasLast = 0;
  sgz_o_break lastBreak;
  memset(&lastBreak, 0, sizeof lastBreak);

  int winLen = halfWinLen * 2;
  if (winLen <= 0) { free(prio.e); return -2; }
  float **eInBuf = alloc_planar(numCh, winLen);

  int afStart = 0;
  if (cfg->hasStart) { afStart = full_to_feat(cfg->spanStart, step); if (afStart < 0) afStart = 0; }
  int afStop = (int)nFrames;
  if (cfg->hasStop) { int s = full_to_feat(cfg->spanStop, step); afStop = s < (int)nFrames ? s : (int)nFrames; }
  int afLen = afStop - afStart;

  int64_t filePos = afStart;
  int left = afLen, readSz = winLen, readOff = 0, logicalOff = 0;
  while (left > 0) {               /* :107-133 */
    int chunkLen = left < readSz ? left : readSz;
    read_frames(file, numCh, filePos, eInBuf, readOff, chunkLen);
    filePos += chunkLen;
    int eInBufOff = logicalOff % winLen;
    sgz_o_normalize(cfg->norm, eInBuf, numCh, readOff, chunkLen);
    float temporal = (tempWeight > 0.0f)
        ? sgz_o_correlate_half(1, halfWinLen, (const float *const *)eInBuf, eInBufOff, 0) : 0.0f;
    float spectral = (tempWeight < 1.0f)
        ? sgz_o_correlate_half(numCoeffs, halfWinLen, (const float *const *)eInBuf, eInBufOff, 1) : 0.0f;
    float sim = temporal * tempWeight + spectral * (1.0f - tempWeight);
    if (curve && logicalOff < curveCap) curve[logicalOff] = sim;
    int hasSpace = prio.n < cfg->numBreaks;                          /* :58 */
    float highest = prio.n > 0 ? prio.e[prio.n - 1].sim : 0.0f;     /* :60-62 */
    if (hasSpace || sim < highest) {
      sgz_o_break b;
      b.sim = sim;
      b._pad = 0;
      b.pos = feat_to_full(afStart + logicalOff + halfWinLen, step);
      /* addBreak, :68-83 */
      if (hasLast && (b.pos - lastBreak.pos) < cfg->minSpacing) {
        if (lastBreak.sim > b.sim) {
          bset_remove(&prio, lastBreak.sim);
          bset_add(&prio, &b);
          lastBreak = b;
        }
      } else {
        bset_add(&prio, &b);
        if (prio.n > cfg->numBreaks) prio.n--;
        lastBreak = b;
        hasLast = 1;
      }
    }
    left -= chunkLen;
    readOff = (readOff + chunkLen) % winLen;
    logicalOff += 1;
    readSz = 1;
  }
  int n = prio.n < cap ? prio.n : cap;
  for (int i = 0; i < n; i++) out[i] = prio.e[i];
  free(prio.e);
  free_planar(eInBuf, numCh);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * SelfSimilarityImpl.body(), Impl/SelfSimilarityImpl.scala:31-180
 * ---------------------------------------------------------------------------------------- */

int sgz_o_self_geometry(const sgz_o_self_cfg *cfg, int64_t nFrames1, int64_t nFrames2,
                        int32_t *decimOut, int32_t *numCorrsOut, int32_t *afStartOut) {
  int step = cfg->stepSize;
  int halfWinLen = full_to_feat(cfg->corrLen, step);
  int winLen = halfWinLen * 2;
  int64_t afNumFrames = nFrames1 < nFrames2 ? nFrames1 : nFrames2;   /* :64 */
  int afStart = 0;
  if (cfg->hasStart) { afStart = full_to_feat(cfg->spanStart, step); if (afStart < 0) afStart = 0; }
  int afStop = (int)afNumFrames;
  if (cfg->hasStop) { int s = full_to_feat(cfg->spanStop, step); afStop = s < (int)afNumFrames ? s : (int)afNumFrames; }
  int afLen = afStop - afStart;
  int64_t n = (int64_t)afLen - winLen + 1;
  if (n < 0) n = 0;
  int numCorrs = (int)n;
  int d = cfg->decimation;
  if (d < 1) return -2;
  int i = numCorrs / d, decim = d, imgExt = i;
  if (i > 0xB504) {                /* :85-90 */
    decim = (numCorrs + 0xB503) / 0xB504;
    imgExt = numCorrs / decim;
  }
  if (decimOut) *decimOut = decim;
  if (numCorrsOut) *numCorrsOut = numCorrs;
  if (afStartOut) *afStartOut = afStart;
  return imgExt;
}

/* colorFun for GrayScale, SelfSimilarityImpl.scala:99-107, applied as in :150 */
static int32_t self_color(const sgz_o_self_cfg *cfg, float sim) {
  float colorScale = 1.0f / cfg->colorCeil;                     /* :115 */
  float m = sim;
  if (!(0.0f >= sim)) m = sim; else m = 0.0f;                   /* math.max(0f, sim) ... */
  if (sim != sim) m = sim;                                      /* ... which propagates NaN */
  float v = (float)pow((double)m, (double)cfg->colorWarp) * colorScale;
  float s = cfg->colorInv ? (1.0f - v) : v;
  float f255 = s * 255;                                         /* Float * Int -> Float */
  int32_t i = j_d2i((double)f255 + 0.5);
  if (i > 255) i = 255;
  if (i < 0) i = 0;
  return (i << 16) | (i << 8) | i;
}

static float self_cell(const sgz_o_self_cfg *cfg, int halfWinLen, float **eInBuf, const float *file1,
                       const float *file2, int afStart, int leftOff, int rightOff) {
  int numCh = cfg->numCh, numCoeffs = numCh - 1;
  float tempWeight = cfg->temporalWeight;
  read_frames(file1, numCh, (int64_t)leftOff + afStart, eInBuf, 0, halfWinLen);            /* :131-133 */
  sgz_o_normalize(cfg->norm, eInBuf, numCh, 0, halfWinLen);
  read_frames(file2, numCh, (int64_t)rightOff + afStart, eInBuf, halfWinLen, halfWinLen);  /* :139-141 */
  sgz_o_normalize(cfg->norm, eInBuf, numCh, halfWinLen, halfWinLen);
  float temporal = (tempWeight > 0.0f)
      ? sgz_o_correlate_half(1, halfWinLen, (const float *const *)eInBuf, 0, 0) : 0.0f;
  float spectral = (tempWeight < 1.0f)
      ? sgz_o_correlate_half(numCoeffs, halfWinLen, (const float *const *)eInBuf, 0, 1) : 0.0f;
  return temporal * tempWeight + spectral * (1.0f - tempWeight);
}

int sgz_o_self_image(const sgz_o_self_cfg *cfg, const float *file1, int64_t nFrames1,
                     const float *file2, int64_t nFrames2, int32_t *rgb, int64_t cap) {
  int32_t decim, numCorrs, afStart;
  int imgExt = sgz_o_self_geometry(cfg, nFrames1, nFrames2, &decim, &numCorrs, &afStart);
  if (imgExt < 0) return imgExt;
  if ((int64_t)imgExt * imgExt > cap) return -3;
  if (!(cfg->colorWarp > 0) || !(cfg->colorCeil > 0)) return -2;
  int halfWinLen = full_to_feat(cfg->corrLen, cfg->stepSize);
  float **eInBuf = alloc_planar(cfg->numCh, halfWinLen * 2);
  int imgExtM1 = imgExt - 1;
  int stop = numCorrs / decim * decim;
  memset(rgb, 0, sizeof(int32_t) * (size_t)imgExt * (size_t)imgExt);
  for (int leftOff = 0; leftOff < stop; leftOff += decim) {
    for (int rightOff = leftOff; rightOff < stop; rightOff += decim) {
      float sim = self_cell(cfg, halfWinLen, eInBuf, file1, file2, afStart, leftOff, rightOff);
      int32_t colr = self_color(cfg, sim);
      int64_t off1 = (int64_t)(imgExtM1 - rightOff / decim) * imgExt + (leftOff / decim);
      int64_t off2 = (int64_t)(imgExtM1 - leftOff / decim) * imgExt + (rightOff / decim);
      rgb[off1] = colr;
      rgb[off2] = colr;
    }
  }
  free_planar(eInBuf, cfg->numCh);
  return imgExt;
}

int sgz_o_self_cells(const sgz_o_self_cfg *cfg, const float *file1, int64_t nFrames1,
                     const float *file2, int64_t nFrames2, int64_t nCells, const int32_t *leftIdx,
                     const int32_t *rightIdx, float *simOut, int32_t *rgbOut) {
  int32_t decim, numCorrs, afStart;
  int imgExt = sgz_o_self_geometry(cfg, nFrames1, nFrames2, &decim, &numCorrs, &afStart);
  if (imgExt < 0) return imgExt;
  int halfWinLen = full_to_feat(cfg->corrLen, cfg->stepSize);
  float **eInBuf = alloc_planar(cfg->numCh, halfWinLen * 2);
  for (int64_t k = 0; k < nCells; k++) {
    int l = leftIdx[k], r = rightIdx[k];
    if (l < 0 || r < 0 || l >= imgExt || r >= imgExt) { free_planar(eInBuf, cfg->numCh); return -2; }
    float sim = self_cell(cfg, halfWinLen, eInBuf, file1, file2, afStart, l * decim, r * decim);
    if (simOut) simOut[k] = sim;
    if (rgbOut) rgbOut[k] = self_color(cfg, sim);
  }
  free_planar(eInBuf, cfg->numCh);
  return imgExt;
}

/* ------------------------------------------------------------------------------------------
 * CrossSimilarityImpl.body(), Impl/CrossSimilarityImpl.scala:32-187
 * ---------------------------------------------------------------------------------------- */

/* openInput, :69-82 -> start frame and length of the span inside a file of numFrames */
static void cross_open(int step, int hasStart, int hasStop, int64_t spStart, int64_t spStop, int64_t numFrames,
                       int64_t *start_out, int64_t *len_out) {
  int64_t s = hasStart ? full_to_feat(spStart, step) : 0;
  int64_t e = hasStop ? full_to_feat(spStop, step) : numFrames;
  int64_t stop = numFrames < e ? numFrames : e;
  int64_t start = stop < s ? stop : s;
  if (start < 0) start = 0;
  *start_out = start;
  *len_out = stop - start;
}

int64_t sgz_o_cross_run(const sgz_o_cross_cfg *cfg, const float *file1, int64_t nFrames1,
                        const float *file2, int64_t nFrames2, float *sim_out, int64_t cap) {
  const int step = cfg->stepSize, numCh = cfg->numCh;
  int64_t st1, l1, st2, l2;
  cross_open(step, cfg->has1Start, cfg->has1Stop, cfg->span1Start, cfg->span1Stop, nFrames1, &st1, &l1);
  cross_open(step, cfg->has2Start, cfg->has2Stop, cfg->span2Start, cfg->span2Stop, nFrames2, &st2, &l2);
  /* shorter file -> afIn1 (read completely), longer -> afIn2 (piecewise), :93-95 */
  const float *fIn1, *fIn2;
  int64_t pos1, len1, pos2, len2;
  if (l1 < l2) { fIn1 = file1; pos1 = st1; len1 = l1; fIn2 = file2; pos2 = st2; len2 = l2; }
  else         { fIn1 = file2; pos1 = st2; len1 = l2; fIn2 = file1; pos2 = st1; len2 = l1; }
  const int bufSz = 8192;
  if (len1 > bufSz) return -2;          /* ArrayIndexOutOfBoundsException in the reference */
  if (len1 <= 0) return len2 > 0 ? -3 : 0;   /* `% len1i` -> ArithmeticException; nothing to do if both are empty */
  const int len1i = (int)len1;

  input_matrix in;                      /* matrixIn, :100-116 */
  memset(&in, 0, sizeof in);
  in.all = alloc_planar(numCh, len1i);
  read_frames(fIn1, numCh, pos1, in.all, 0, len1i);
  sgz_o_normalize(cfg->norm, in.all, numCh, 0, len1i);
  in.numFrames = len1i;
  in.temporal.mat = in.all; in.temporal.numChannels = 1; in.temporal.numFrames = len1i;
  sgz_o_stat((const float *const *)in.temporal.mat, 0, len1i, 0, 1, &in.temporal.mean, &in.temporal.stdDev);
  in.spectral.mat = in.all + 1; in.spectral.numChannels = numCh - 1; in.spectral.numFrames = len1i;
  sgz_o_stat((const float *const *)in.spectral.mat, 0, len1i, 0, numCh - 1, &in.spectral.mean, &in.spectral.stdDev);
  in.lnAvgLoudness = calc_ln_avg_loud(in.all[0], 0, len1i);

  const float w = cfg->temporalWeight;
  float **eInBuf = alloc_planar(numCh, bufSz);
  int64_t left = len2, filePos = pos2, nOut = 0;
  int readSz = bufSz, readOff = 0, logicalOff = 0;
  while (left > 0) {                    /* :135-171 */
    int chunkLen = (int)(left < readSz ? left : readSz);
    read_frames(fIn2, numCh, filePos, eInBuf, readOff, chunkLen);
    filePos += chunkLen;
    int eInBufOff = logicalOff % len1i;
    sgz_o_normalize(cfg->norm, eInBuf, numCh, readOff, chunkLen);
    float boost = calc_boost(&in, eInBuf[0]);
    float sim;
    if (boost <= cfg->maxBoost) {
      /* private correlate (:177-185): stat over b[0, numFrames), MathUtil.correlate wraps % b.length = 8192 */
      float temporal = (w > 0.0f) ? corr_wrap(&in.temporal, (const float *const *)eInBuf, bufSz, eInBufOff, 0) : 0.0f;
      float spectral = (w < 1.0f) ? corr_wrap(&in.spectral, (const float *const *)eInBuf, bufSz, eInBufOff, 1) : 0.0f;
      sim = temporal * w + spectral * (1.0f - w);
    } else {
      sim = 0.0f;
    }
    if (nOut < cap) sim_out[nOut] = sim;
    nOut++;
    left -= chunkLen;
    readOff = (readOff + chunkLen) % len1i;
    logicalOff += 1;
    readSz = 1;
  }
  free_planar(eInBuf, numCh);
  free_planar(in.all, numCh);
  return nOut;
}

/* ------------------------------------------------------------------------------------------
 * FeatureStatsImpl, Impl/FeatureStatsImpl.scala:30-135
 * ---------------------------------------------------------------------------------------- */

static double j_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }   /* math.min */
static double j_max(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }   /* math.max */

/* body1, :55-135: (p01[ch], p99[ch]) of one file */
static void stats_body1(int numCh, const float *file, int64_t numFrames, double *p01, double *p99) {
  float *maxs = (float *)malloc(sizeof(float) * (size_t)numCh);
  float *mins = (float *)malloc(sizeof(float) * (size_t)numCh);
  double *sums = (double *)calloc((size_t)numCh, sizeof(double));
  double *skews = (double *)calloc((size_t)numCh, sizeof(double));
  int32_t *pctils = (int32_t *)calloc((size_t)numCh * 2048, sizeof(int32_t));
  for (int ch = 0; ch < numCh; ch++) { maxs[ch] = -INFINITY; mins[ch] = INFINITY; }
  /* first pass (:70-84); the 8192-frame chunking does not change the per-channel order */
  for (int ch = 0; ch < numCh; ch++) {
    for (int64_t i = 0; i < numFrames; i++) {
      float f = file[i * numCh + ch];
      if (f < mins[ch]) mins[ch] = f;
      if (f > maxs[ch]) maxs[ch] = f;
      sums[ch] += f;
    }
  }
  const double log05 = log(0.5);
  for (int ch = 0; ch < numCh; ch++) {       /* :86-92 */
    double mean = sums[ch] / (double)numFrames;
    float d = maxs[ch] - mins[ch];
    double mn = (mean - mins[ch]) / d;
    skews[ch] = log05 / log(mn);
  }
  for (int ch = 0; ch < numCh; ch++) {       /* second pass, :94-113 */
    int32_t *cp = pctils + (size_t)ch * 2048;
    float mn = mins[ch];
    float d = maxs[ch] - mn;
    double skew = skews[ch];
    for (int64_t i = 0; i < numFrames; i++) {
      float f = file[i * numCh + ch];
      int32_t norm = j_d2i(pow((double)((f - mn) / d), skew) * 2047 + 0.5);
      if (norm >= 0 && norm < 2048) cp[norm] += 1;   /* outside: ArrayIndexOutOfBounds upstream; cannot happen for finite data */
    }
  }
  for (int ch = 0; ch < numCh; ch++) {       /* :115-131 */
    const int32_t *cp = pctils + (size_t)ch * 2048;
    int32_t p01n = j_d2i((double)numFrames * 0.01);
    int32_t p99n = j_d2i((double)numFrames * 0.99);
    double skewr = 1.0 / skews[ch];
    float mn = mins[ch];
    float d = maxs[ch] - mn;
    int32_t cnt = 0;
    int i = 0;
    while (cnt < p01n && i < 2048) { cnt += cp[i]; i++; }
    p01[ch] = pow((double)i / 2048, skewr) * d + mn;
    while (cnt < p99n && i < 2048) { cnt += cp[i]; i++; }
    p99[ch] = pow((double)i / 2048, skewr) * d + mn;
  }
  free(maxs); free(mins); free(sums); free(skews); free(pctils);
}

int sgz_o_stats_run(int numCh, int numFiles, const float *const *files, const int64_t *nFrames, double *out,
                    double *perFile) {
  double *p01 = (double *)malloc(sizeof(double) * (size_t)numCh);
  double *p99 = (double *)malloc(sizeof(double) * (size_t)numCh);
  for (int f = 0; f < numFiles; f++) {       /* body, :30-53 */
    stats_body1(numCh, files[f], nFrames[f], p01, p99);
    for (int ch = 0; ch < numCh; ch++) {
      if (perFile) { perFile[((size_t)f * numCh + ch) * 2] = p01[ch]; perFile[((size_t)f * numCh + ch) * 2 + 1] = p99[ch]; }
      if (f == 0) { out[2 * ch] = p01[ch]; out[2 * ch + 1] = p99[ch]; }
      else { out[2 * ch] = j_min(out[2 * ch], p01[ch]); out[2 * ch + 1] = j_max(out[2 * ch + 1], p99[ch]); }
    }
  }
  free(p01); free(p99);
  return numFiles;
}
