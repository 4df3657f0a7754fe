"""The oracle's match queues against a second, independently written restatement (tests/queue_model.py): both replay
FeatureCorrelationImpl.body() (FeatureCorrelationImpl.scala:113-150, 190-246, 250-400) on the same similarity curves and
must return the same matches bit for bit -- over the randomly drawn configurations of the GPU differential tests (ragged and
too-short files, repeated files = equal sims, silence = NaN sims, quiet files = gated sims, any numMatches / numPerFile /
minSpacing, punch-in only and punch-in + punch-out)."""
import numpy as np
import pytest

import queue_model
from test_gpu_fuzz import draw_case
from util import O, STEP, corr_cfgs, make_db, make_input, synth


def same(model, oracle):
    assert len(model) == len(oracle), (len(model), len(oracle))
    for i, (a, b) in enumerate(zip(model, oracle)):
        for key in ("file", "start", "stop"):
            assert a[key] == b[key], (i, key, a, b)
        for key in ("sim", "boostIn", "boostOut"):
            x, y = np.float32(a[key]), np.float32(b[key])
            assert x.tobytes() == y.tobytes() or (np.isnan(x) and np.isnan(y)), (i, key, a, b)


@pytest.mark.parametrize("block", range(8))
def test_punch_in_queues(block):
    for seed in range(15 * block, 15 * block + 15):
        inp, files, norm, cfg = draw_case(seed, False)
        op, _ = corr_cfgs(inp, norm, **cfg)
        same(queue_model.search(O, op, files), O.corr_search(op, files))


@pytest.mark.parametrize("block", range(8))
def test_punch_out_queues(block):
    n = 0
    for seed in range(200 + 10 * block, 200 + 10 * block + 10):
        inp, files, norm, cfg = draw_case(seed, True)
        op, _ = corr_cfgs(inp, norm, **cfg)
        want = O.corr_search(op, files)
        same(queue_model.search(O, op, files), want)
        n += len(want)
    assert n > 0


def test_repeated_files_silence_and_spacing():
    """equal sims (a file three times), NaN sims (a constant stretch) and a tight minSpacing in one search, with and
    without punch-out"""
    files, norm = make_db(6, [900, 700, 900, 650, 900, 300])
    inp = make_input(700)
    files[2] = files[0].copy()
    files[4] = files[0].copy()
    files[1][100:420] = files[1][100]
    files[3][50:222] = synth.plant(inp[:172], 5, 1)
    for po in (None, (345 * STEP, 517 * STEP)):
        for nm, npf, sp in ((4, 2, 0), (10, 3, 30 * STEP), (3, 1, 10 ** 7), (50, 50, 0)):
            op, _ = corr_cfgs(inp, norm, punch_out=po, num_matches=nm, num_per_file=npf, min_spacing=sp,
                              min_punch=40 * STEP, max_punch=300 * STEP)
            same(queue_model.search(O, op, files), O.corr_search(op, files))


@pytest.mark.parametrize("seed", range(24))
def test_segmentation_break_queue(seed):
    """FeatureSegmentationImpl's break queue (:52-82) restated on the oracle's curve: same breaks bit for bit"""
    rng = np.random.default_rng(400 + seed)
    num_ch = int(rng.choice([3, 14]))
    _, _, _, norm = synth.default_profile(num_ch)
    n = int(rng.integers(40, 1500))
    if seed % 3 == 0:
        f, _ = synth.regime_file(synth.BASE_SEED, 30 + seed, n, num_ch, int(rng.integers(2, 9)))
    else:
        mu, sigma, floor0, _ = synth.default_profile(num_ch)
        f = synth.synth_file(synth.BASE_SEED, 30 + seed, n, mu, sigma, floor0)
    if seed % 5 == 1 and n > 300:
        f[100:260] = f[100]                               # silence: NaN sims
    half = int(rng.choice([1, 2, 7, 43, 86]))
    use_span = bool(rng.integers(0, 2))
    sp = O.SegmParams(step_size=STEP, corr_len=half * STEP + int(rng.integers(-200, 200)), temporal_weight=float(rng.choice([0, 1, 0.5, 0.3])),
                      norm=norm if rng.integers(0, 3) else None, num_breaks=int(rng.choice([1, 2, 5, 20, 100])),
                      min_spacing=int(rng.choice([0, 1, 512, 22050, 10 ** 6])),
                      span_start=int(rng.integers(0, n // 3)) * STEP + 17 if use_span else None,
                      span_stop=int(rng.integers(2 * n // 3, n + 50)) * STEP - 23 if use_span else None)
    want = O.segm_run(sp, f)
    got = queue_model.segmentation(O, sp, f)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a["pos"] == b["pos"], (a, b)
        x, y = np.float32(a["sim"]), np.float32(b["sim"])
        assert x.tobytes() == y.tobytes() or (np.isnan(x) and np.isnan(y)), (a, b)


@pytest.mark.parametrize("num_ch,w,weight,use_norm,max_boost", [
    (14, 43, 0.5, True, 8.0), (5, 16, 0.3, True, 8.0), (2, 3, 1.0, False, 8.0), (3, 7, 0.0, True, 1.05), (14, 20, 0.7, False, 2.0),
])
def test_offset_arithmetic_restated(num_ch, w, weight, use_norm, max_boost):
    """tests/arith_model.py (normalize, stat in ring order, correlate in logical order, calcBoost, the Float blend) gives the
    oracle's similarity curve BIT FOR BIT, silence (0 / 0 = NaN) and gated offsets (boost > maxBoost -> 0f) included"""
    import warnings

    import arith_model
    mu, sigma, floor0, norm = synth.default_profile(num_ch)
    inp = synth.synth_file(synth.BASE_SEED, 0, w + 9, mu, sigma, floor0)
    f = synth.synth_file(synth.BASE_SEED, 77, 140, mu, sigma, floor0)
    f[60:60 + w + 5] = f[60]                                   # a constant stretch: NaN sims
    f[115:, 0] *= np.float32(0.6)                              # a quieter stretch: boosts above a tight maxBoost
    norm = norm if use_norm else None
    op, _ = corr_cfgs(inp, norm, punch_in=(4 * STEP, (4 + w) * STEP), w_in=weight, max_boost=max_boost)
    want_sim, want_boost = O.corr_curve(op, f, 0, 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                        # 0 / 0 in numpy scalars warns
        got_sim, got_boost = arith_model.curve(inp, 4, w, norm, weight, max_boost, f)
    assert len(got_sim) == len(want_sim) == 140 - w + 1
    a, b = np.array(got_sim, np.float32), np.asarray(want_sim, np.float32)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and (np.isnan(b).any() or weight == 0.0)
    assert np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))
    assert np.array_equal(np.array(got_boost, np.float32).view(np.uint32), np.asarray(want_boost, np.float32).view(np.uint32))
    if max_boost < 2:
        assert (b == 0).any()


@pytest.mark.parametrize("num_ch,half,weight,use_norm,n,span", [
    (14, 21, 0.5, True, 130, None), (3, 2, 1.0, False, 60, (5, 50)), (5, 8, 0.0, True, 90, None), (14, 43, 0.3, True, 70, (0, 70)),
])
def test_segmentation_arithmetic_restated(num_ch, half, weight, use_norm, n, span):
    """tests/arith_model.py: correlateHalf over the rotating buffer gives the oracle's segmentation curve bit for bit (a
    span shorter than the window -- the last case -- yields the one offset on the zero-filled rest of the buffer)"""
    import warnings

    import arith_model
    mu, sigma, floor0, norm = synth.default_profile(num_ch)
    f = synth.synth_file(synth.BASE_SEED, 55, n, mu, sigma, floor0)
    if n > 100:
        f[40:40 + 2 * half + 4] = f[40]                        # silence: 0 / 0
    norm = norm if use_norm else None
    sp = O.SegmParams(step_size=STEP, corr_len=half * STEP, temporal_weight=weight, norm=norm, num_breaks=3, min_spacing=0,
                      span_start=None if span is None else span[0] * STEP, span_stop=None if span is None else span[1] * STEP)
    _, want = O.segm_run(sp, f, want_curve=True)
    a0, a1 = (0, n) if span is None else span
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = np.array(arith_model.segm_curve(f, half, norm, weight, a0, a1), np.float32)
    want = np.asarray(want[:len(got)], np.float32)
    assert len(got) == (a1 - a0 - 2 * half + 1 if a1 - a0 >= 2 * half else 1)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))


@pytest.mark.parametrize("num_ch,half,decim,weight,inv,warp,ceil,n", [
    (14, 6, 1, 0.5, False, 1.0, 1.0, 60), (5, 3, 2, 0.3, True, 0.5, 0.8, 50), (3, 4, 3, 1.0, False, 2.0, 1.0, 47), (14, 5, 1, 0.0, True, 1.0, 2.0, 40),
])
def test_self_similarity_restated(num_ch, half, decim, weight, inv, warp, ceil, n):
    """tests/arith_model.py: SelfSimilarityImpl's cell loop, decimation, mirror pixels and GrayScale colour function
    (:76-150) give the oracle's image pixel for pixel"""
    import warnings

    import arith_model
    mu, sigma, floor0, norm = synth.default_profile(num_ch)
    f = synth.synth_file(synth.BASE_SEED, 66, n, mu, sigma, floor0)
    f[20:20 + 2 * half + 3] = f[20]                            # silence: NaN sims -> (NaN * 255 + 0.5).toInt = 0
    p = O.SelfParams(step_size=STEP, corr_len=half * STEP, decimation=decim, temporal_weight=weight, norm=norm,
                     color_inv=inv, color_warp=warp, color_ceil=ceil)
    want = O.self_image(p, f)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = np.array(arith_model.self_image(f, f, half, decim, norm, weight, inv, warp, ceil, 0, n), np.int64)
    assert got.shape == want.shape and got.shape[0] == (n - 2 * half + 1) // decim
    assert np.array_equal(got, want.astype(np.int64) & 0xFFFFFF)


def test_feature_stats_restated():
    """tests/arith_model.py: FeatureStatsImpl's two passes (min / max / mean, skewed 2048-bin histogram, percentiles) give the
    oracle's normalisation ranges to the last bit of the Doubles"""
    import arith_model
    mu, sigma, floor0, _ = synth.default_profile(5)
    files = [synth.synth_file(synth.BASE_SEED, 90 + i, n, mu, sigma, floor0) for i, n in enumerate((700, 1500, 333))]
    files[1][:, 2] *= np.float32(3.0)
    want, per = O.stats_run(files, want_per_file=True)
    mins, maxs, per_file = arith_model.feature_stats(files)
    assert np.array_equal(np.array(mins), want[:, 0]) and np.array_equal(np.array(maxs), want[:, 1])
    for i, (p01, p99) in enumerate(per_file):
        assert np.array_equal(np.array(p01), per[i, :, 0]) and np.array_equal(np.array(p99), per[i, :, 1])


@pytest.mark.parametrize("n1,n2,num_ch,weight", [(25, 8192 + 120, 5, 0.5), (8192 + 60, 17, 3, 0.3), (40, 300, 14, 1.0)])
def test_cross_similarity_restated(n1, n2, num_ch, weight):
    """tests/arith_model.py: CrossSimilarityImpl's loop with its 8192-frame buffer quirks gives the oracle's curve bit for bit"""
    import warnings

    import arith_model
    mu, sigma, floor0, norm = synth.default_profile(num_ch)
    f1 = synth.synth_file(synth.BASE_SEED, 41, n1, mu, sigma, floor0)
    f2 = synth.synth_file(synth.BASE_SEED, 42, n2, mu, sigma, floor0)
    want = O.cross_run(O.CrossParams(step_size=STEP, norm=norm, temporal_weight=weight), f1, f2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = np.array(arith_model.cross_curve(f1, f2, norm, weight, 8.0), np.float32)
    assert got.shape == want.shape == (1 + max(n1, n2) - min(max(n1, n2), 8192),)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32))
