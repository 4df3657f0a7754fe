"""The arithmetic of the five processes, restated a second time -- test infrastructure, like the oracle
(FeatureCorrelation's offset, FeatureSegmentation's correlateHalf, the SelfSimilarity image, CrossSimilarity, FeatureStats).

Pure-Python loops over IEEE doubles (Python's float) and numpy.float32 scalars, written from MathUtil.scala (stat :29-62, avg
:109-118, normalize :132-152, correlate :177-196) and FeatureCorrelationImpl.scala (readInBuffer :83-98, calcBoost :73-78, the
loop body :190-211): same conversion points, same order of every addition (stat walks the ring buffer in PHYSICAL order,
correlate in logical order), Float arithmetic where the Scala has Floats.  tests/test_queue_model.py compares the curves it
produces with the oracle's bit for bit on small cases; together with queue_model.py (the queues) every part of the oracle's
FeatureCorrelation has an independently written twin."""
import math

import numpy as np

F32 = np.float32


def normalize(norm, b):
    """MathUtil.normalize: Float arithmetic, in place; b = [ch][frames] of np.float32"""
    if norm is None:
        return
    for ch in range(len(b)):
        mn, mx = F32(norm[ch][0]), F32(norm[ch][1])
        d = F32(mx - mn)
        for i in range(len(b[ch])):
            b[ch][i] = F32(F32(b[ch][i] - mn) / d)


def stat(mat, frame_len, chan_off, chan_len):
    """MathUtil.stat over ring positions 0 .. frame_len-1 of rows chan_off .. : (mean, stddev) in Double"""
    s = 0.0
    for ch in range(chan_off, chan_off + chan_len):
        for i in range(frame_len):
            s += float(mat[ch][i])
    size = frame_len * chan_len
    mean = s / size
    s = 0.0
    for ch in range(chan_off, chan_off + chan_len):
        for i in range(frame_len):
            d = float(mat[ch][i]) - mean
            s += d * d
    return mean, math.sqrt(s / size)


def avg(row, n):
    s = 0.0
    for i in range(n):
        s += float(row[i])
    return F32(s / n)


def correlate(a, a_mean, a_std, num_frames, num_ch, b, b_mean, b_std, b_frame_off, b_chan_off):
    a_add, b_add = -a_mean, -b_mean
    s = 0.0
    for ch in range(num_ch):
        ca, cb = a[ch], b[ch + b_chan_off]
        for i in range(num_frames):
            s += (float(ca[i]) + a_add) * (float(cb[(i + b_frame_off) % len(cb)]) + b_add)
    # (IEEE division: 0 / 0 = NaN for a constant window, where Python's float raises)
    return F32(np.float64(s) / np.float64(a_std * b_std * (num_ch * num_frames)))


def curve(inp, start_frame, w, norm, weight, max_boost, file):
    """(sim, boost) of every offset of `file` for the punch window inp[start_frame : start_frame + w] (loop A, no gate by
    the queues): inp, file = [frames][ch] float32 arrays"""
    num_ch = inp.shape[1]
    a = [[F32(inp[start_frame + i][ch]) for i in range(w)] for ch in range(num_ch)]
    normalize(norm, a)
    t_mean, t_std = stat(a, w, 0, 1)
    s_mean, s_std = stat(a, w, 1, num_ch - 1)
    ln_avg_in = math.log(float(avg(a[0], w)))
    wt = F32(weight)
    n = file.shape[0]
    ring = [[F32(0.0)] * w for _ in range(num_ch)]
    sims, boosts = [], []
    left, read_sz, read_off, logical = n, w, 0, 0
    pos = 0
    if left < w:
        return sims, boosts                       # (no offset)
    while left > 0:
        chunk = min(left, read_sz)
        fresh = [[F32(file[pos + k][ch]) for k in range(chunk)] for ch in range(num_ch)]
        normalize(norm, fresh)
        for ch in range(num_ch):
            for k in range(chunk):
                ring[ch][read_off + k] = fresh[ch][k]
        pos += chunk
        off = logical % w
        boost = F32(math.exp((ln_avg_in - math.log(float(avg(ring[0], w)))) / 0.6))
        if boost <= F32(max_boost):
            if wt > 0:
                b_mean, b_std = stat(ring, w, 0, 1)
                temporal = correlate(a[0:1], t_mean, t_std, w, 1, ring, b_mean, b_std, off, 0)
            else:
                temporal = F32(0.0)
            if wt < 1:
                b_mean, b_std = stat(ring, w, 1, num_ch - 1)
                spectral = correlate(a[1:], s_mean, s_std, w, num_ch - 1, ring, b_mean, b_std, off, 1)
            else:
                spectral = F32(0.0)
            sim = F32(F32(temporal * wt) + F32(spectral * F32(F32(1.0) - wt)))
        else:
            sim = F32(0.0)
        sims.append(sim)
        boosts.append(boost)
        left -= chunk
        read_off = (read_off + chunk) % w
        logical += 1
        read_sz = 1
    return sims, boosts


def correlate_half(num_ch, half, a, frame_off, chan_off):
    """MathUtil.correlateHalf (:80-99): the two halves of the ring buffer against each other"""
    num_frames = half << 1
    mean, std = stat(a, num_frames, chan_off, num_ch)
    add = -mean
    s = 0.0
    for ch in range(num_ch):
        ca = a[ch + chan_off]
        for i in range(frame_off, frame_off + half):
            s += (float(ca[i % num_frames]) + add) * (float(ca[(i + half) % num_frames]) + add)
    return F32(np.float64(s) / np.float64(std * std * (num_ch * half)))


def segm_curve(file, half, norm, weight, af_start, af_stop):
    """the similarity of every offset of FeatureSegmentationImpl.body()'s loop (:103-127): file = [frames][ch] float32"""
    num_ch = file.shape[1]
    win = 2 * half
    wt = F32(weight)
    ring = [[F32(0.0)] * win for _ in range(num_ch)]      # eInBuf: freshly allocated, zero-filled
    left, read_sz, read_off, logical, pos = af_stop - af_start, win, 0, 0, af_start
    sims = []
    while left > 0:
        chunk = min(left, read_sz)
        fresh = [[F32(file[pos + k][ch]) for k in range(chunk)] for ch in range(num_ch)]
        normalize(norm, fresh)
        for ch in range(num_ch):
            for k in range(chunk):
                ring[ch][read_off + k] = fresh[ch][k]
        pos += chunk
        off = logical % win
        temporal = correlate_half(1, half, ring, off, 0) if wt > 0 else F32(0.0)
        spectral = correlate_half(num_ch - 1, half, ring, off, 1) if wt < 1 else F32(0.0)
        sims.append(F32(F32(temporal * wt) + F32(spectral * F32(F32(1.0) - wt))))
        left -= chunk
        read_off = (read_off + chunk) % win
        logical += 1
        read_sz = 1
    return sims


def self_image(file1, file2, half, decimation, norm, weight, color_inv, color_warp, color_ceil, af_start, af_stop):
    """SelfSimilarityImpl.body() (:64-167), GrayScale: the image as a list of rows of 0xRRGGBB ints"""
    num_ch = file1.shape[1]
    win = 2 * half
    af_len = af_stop - af_start
    num_corrs = max(0, af_len - win + 1)
    i = num_corrs // decimation
    if i <= 0xB504:
        decim, ext = decimation, i
    else:
        decim = (num_corrs + 0xB503) // 0xB504
        ext = num_corrs // decim
    wt = F32(weight)
    scale = F32(F32(1.0) / F32(color_ceil))
    img = [[0] * ext for _ in range(ext)]
    stop = num_corrs // decim * decim
    buf = [[F32(0.0)] * win for _ in range(num_ch)]

    def read(f, at, dst):
        fresh = [[F32(f[at + k][ch]) for k in range(half)] for ch in range(num_ch)]
        normalize(norm, fresh)
        for ch in range(num_ch):
            for k in range(half):
                buf[ch][dst + k] = fresh[ch][k]

    left = 0
    while left < stop:
        read(file1, left + af_start, 0)
        right = left
        while right < stop:
            read(file2, right + af_start, half)
            temporal = correlate_half(1, half, buf, 0, 0) if wt > 0 else F32(0.0)
            spectral = correlate_half(num_ch - 1, half, buf, 0, 1) if wt < 1 else F32(0.0)
            sim = F32(F32(temporal * wt) + F32(spectral * F32(F32(1.0) - wt)))
            # colorFun(math.pow(math.max(0f, sim), colorWarp).toFloat * colorScale); math.max(0f, NaN) = NaN
            base = sim if (sim > 0 or np.isnan(sim)) else F32(0.0)
            v = F32(F32(math.pow(float(base), float(F32(color_warp)))) * scale)
            x = F32(F32(1.0) - v) if color_inv else v
            d = float(F32(x * F32(255))) + 0.5                    # Float * Int -> Float, + 0.5 -> Double
            g = 0 if math.isnan(d) else int(max(-2147483648.0, min(2147483647.0, math.trunc(d))))      # Double.toInt
            g = max(0, min(255, g))
            colr = (g << 16) | (g << 8) | g
            img[(ext - 1 - right // decim)][left // decim] = colr
            img[(ext - 1 - left // decim)][right // decim] = colr
            right += decim
        left += decim
    return img


def feature_stats(files):
    """FeatureStatsImpl.body() / body1() (:30-135): per file the 1st and 99th percentile of every channel through a skewed
    2048-bin histogram; over the files the minimum of the former and the maximum of the latter.  files = [frames][ch]"""
    all_mins = all_maxs = None
    per_file = []
    for f in files:
        n, num_ch = f.shape
        mins = [F32(np.inf)] * num_ch
        maxs = [F32(-np.inf)] * num_ch
        sums = [0.0] * num_ch
        for ch in range(num_ch):
            for i in range(n):
                v = F32(f[i][ch])
                if v < mins[ch]:
                    mins[ch] = v
                if v > maxs[ch]:
                    maxs[ch] = v
                sums[ch] += float(v)
        log05 = math.log(0.5)
        p01, p99 = [0.0] * num_ch, [0.0] * num_ch
        for ch in range(num_ch):
            mean = sums[ch] / n
            d = F32(maxs[ch] - mins[ch])                       # Float - Float
            mn = (mean - float(mins[ch])) / float(d)
            skew = log05 / math.log(mn)
            hist = [0] * 2048
            mn_f = mins[ch]
            for i in range(n):
                v = F32(f[i][ch])
                x = float(F32(F32(v - mn_f) / d))              # (f - min) / d in Float, widened for math.pow
                hist[int(math.pow(x, skew) * 2047 + 0.5)] += 1
            p01n, p99n = int(n * 0.01), int(n * 0.99)
            skewr = 1.0 / skew
            cnt = i = 0
            while cnt < p01n:
                cnt += hist[i]
                i += 1
            p01[ch] = math.pow(i / 2048, skewr) * float(d) + float(mn_f)
            while cnt < p99n:
                cnt += hist[i]
                i += 1
            p99[ch] = math.pow(i / 2048, skewr) * float(d) + float(mn_f)
        per_file.append((p01, p99))
        if all_mins is None:
            all_mins, all_maxs = list(p01), list(p99)
        else:
            for ch in range(num_ch):
                all_mins[ch] = min(all_mins[ch], p01[ch])
                all_maxs[ch] = max(all_maxs[ch], p99[ch])
    return all_mins, all_maxs, per_file


def cross_curve(file1, file2, norm, weight, max_boost):
    """CrossSimilarityImpl.body() (:83-176) for Span.All inputs: the shorter file is the template, the longer one runs
    through the 8192-frame buffer -- whose first read takes 8192 frames at once, whose write position wraps at the TEMPLATE
    length and whose read index wraps at the BUFFER length, exactly as the reference has it"""
    f_in1, f_in2 = (file1, file2) if file1.shape[0] < file2.shape[0] else (file2, file1)
    len1, len2, num_ch = f_in1.shape[0], f_in2.shape[0], f_in1.shape[1]
    a = [[F32(f_in1[i][ch]) for i in range(len1)] for ch in range(num_ch)]
    normalize(norm, a)
    t_mean, t_std = stat(a, len1, 0, 1)
    s_mean, s_std = stat(a, len1, 1, num_ch - 1)
    ln_avg_in = math.log(float(avg(a[0], len1)))
    wt = F32(weight)
    buf_sz = 8192
    ring = [[F32(0.0)] * buf_sz for _ in range(num_ch)]
    left, read_sz, read_off, logical, pos = len2, buf_sz, 0, 0, 0
    out = []
    while left > 0:
        chunk = min(left, read_sz)
        fresh = [[F32(f_in2[pos + k][ch]) for k in range(chunk)] for ch in range(num_ch)]
        normalize(norm, fresh)
        for ch in range(num_ch):
            ring[ch][read_off:read_off + chunk] = fresh[ch]
        pos += chunk
        off = logical % len1
        boost = F32(math.exp((ln_avg_in - math.log(float(avg(ring[0], len1)))) / 0.6))
        if boost <= F32(max_boost):
            if wt > 0:
                b_mean, b_std = stat(ring, len1, 0, 1)
                temporal = correlate(a[0:1], t_mean, t_std, len1, 1, ring, b_mean, b_std, off, 0)
            else:
                temporal = F32(0.0)
            if wt < 1:
                b_mean, b_std = stat(ring, len1, 1, num_ch - 1)
                spectral = correlate(a[1:], s_mean, s_std, len1, num_ch - 1, ring, b_mean, b_std, off, 1)
            else:
                spectral = F32(0.0)
            out.append(F32(F32(temporal * wt) + F32(spectral * F32(F32(1.0) - wt))))
        else:
            out.append(F32(0.0))
        left -= chunk
        read_off = (read_off + chunk) % len1
        logical += 1
        read_sz = 1
    return out
