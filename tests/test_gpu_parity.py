"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (north_star): correlation scores within 1e-5 relative; identical match file/offset positions;
bit-identical segmentation breaks; (this round) pixel-identical self-similarity images.
"""
import numpy as np
import pytest

from util import (N, O, STEP, assert_matches_equal, assert_sims_close, build_db, corr_cfgs, make_db, make_input,
                  plant_needles, synth)

pytestmark = pytest.mark.gpu


def test_library_reports_b200(ctx):
    from strugatzki_b200 import engine
    assert engine.device_count() >= 1
    assert N.lib().sgz_abi_version() == 1


def test_synth_device_equals_numpy(ctx):
    """device generator == numpy generator, bit for bit, after the reference's normalisation."""
    from strugatzki_b200 import engine
    mu, sigma, floor0, norm = synth.default_profile(14)
    db = engine.Database(ctx, 14, norm)
    for stream, n in ((3, 1000), (4, 777)):
        db.add_synth(synth.BASE_SEED, stream, n, mu, sigma, float(floor0))
    db.finalize()
    for idx, (stream, n) in enumerate(((3, 1000), (4, 777))):
        raw = synth.synth_file(synth.BASE_SEED, stream, n, mu, sigma, floor0)
        want = ((raw - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32).T
        got = db.read(idx, 0, n)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_db_layouts_and_patch(ctx):
    from strugatzki_b200 import engine
    files, norm = make_db(1, 600)
    f = files[0]
    db = engine.Database(ctx, 14, norm)
    db.add_file(f, N.LAYOUT_INTERLEAVED_LE)
    db.add_file(f.astype(">f4"), N.LAYOUT_INTERLEAVED_BE)
    db.add_file(np.ascontiguousarray(f.T), N.LAYOUT_PLANAR_LE)
    db.add_file(np.zeros((0, 14), np.float32))          # empty file is legal
    patch = make_input(50)
    db.patch(0, 100, patch)
    db.finalize()
    assert db.info() == (4, 1800, 14)
    want = ((f - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32).T
    a, b, c = db.read(0, 0, 600), db.read(1, 0, 600), db.read(2, 0, 600)
    assert np.array_equal(b, want) and np.array_equal(c, want)
    assert np.array_equal(a[:, :100], want[:, :100]) and np.array_equal(a[:, 150:], want[:, 150:])
    wp = ((patch - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32).T
    assert np.array_equal(a[:, 100:150], wp)
    # normalize = false keeps the raw values
    db2 = engine.Database(ctx, 14, None)
    db2.add_file(f)
    db2.finalize()
    assert np.array_equal(db2.read(0, 0, 600), f.T)


@pytest.mark.parametrize("w_in,weight,num_ch", [(88200, 0.5, 14), (88200, 0.0, 14), (88200, 1.0, 14),
                                                  (22050, 0.3, 14), (50000, 0.5, 5), (100000, 0.7, 21)])
def test_corr_curve_matches_oracle(ctx, w_in, weight, num_ch):
    """every punch-in offset: sim within 1e-5 relative, boost within 1e-6 relative"""
    from strugatzki_b200 import engine
    files, norm = make_db(3, [2600, 400, 3100], num_ch)
    inp = make_input(900, num_ch)
    W = (w_in + STEP // 2) // STEP
    plant_needles(files, inp[:W], [(0, 1000), (2, 37)])
    op, nc = corr_cfgs(inp, norm, punch_in=(0, w_in), w_in=weight, num_matches=3)
    db = build_db(ctx, files, norm)
    job = engine.CorrelationJob(db, nc, inp)
    job.scan()
    total = 0
    for i, f in enumerate(files):
        want_sim, want_boost = O.corr_curve(op, f, 0, 0)
        n = len(want_sim)
        assert n == max(0, f.shape[0] - W + 1)
        total += n
        if n == 0:
            continue
        sim, boost = job.curve(i, 0, 0, n)
        assert_sims_close(sim, want_sim, rel=1e-5, abs_tol=2e-6, what=f"file {i} sim")
        assert_sims_close(boost, want_boost, rel=1e-5, abs_tol=0, what=f"file {i} boost")
    assert job.num_offsets == total == O.corr_num_offsets(op, [f.shape[0] for f in files])


@pytest.mark.parametrize("num_matches,num_per_file,min_spacing", [
    (1, 1, 0), (5, 1, 0), (5, 2, 0), (8, 3, 22050), (20, 4, 44100), (40, 1, 0), (7, 100, 0), (3, 2, -10 ** 9),
])
def test_corr_search_matches_oracle(ctx, num_matches, num_per_file, min_spacing):
    """FeatureCorrelation punch-in search: identical files / spans, sims within 1e-5"""
    from strugatzki_b200 import engine
    files, norm = make_db(12, [3000, 2500, 4100, 180, 2900, 3300, 2000, 5000, 172, 2600, 3100, 2800])
    inp = make_input(900)
    W = 172
    plant_needles(files, inp[:W], [(2, 1500), (7, 10), (7, 3000), (9, 2400), (11, 100)])
    op, nc = corr_cfgs(inp, norm, num_matches=num_matches, num_per_file=num_per_file, min_spacing=min_spacing)
    db = build_db(ctx, files, norm)
    job = engine.CorrelationJob(db, nc, inp)
    got = job.run()
    want = O.corr_search(op, files)
    assert_matches_equal(got, want)


def test_corr_search_boost_gate_and_no_norm(ctx):
    from strugatzki_b200 import engine
    files, _ = make_db(5, 2500)
    inp = make_input(900)
    files[1][:, 0] *= np.float32(0.05)        # very quiet file -> boost > maxBoost -> sim forced to 0f
    plant_needles(files, inp[:172], [(3, 700)])
    op, nc = corr_cfgs(inp, None, num_matches=4, num_per_file=1, max_boost=2.0)
    db = build_db(ctx, files, None)
    job = engine.CorrelationJob(db, nc, inp)
    got = job.run()
    want = O.corr_search(op, files)
    assert_matches_equal(got, want)
    sim, boost = job.curve(1, 0, 0, 100)
    assert np.all(sim == 0.0) and np.all(boost > 2.0)


def test_corr_async_poll_and_abort(ctx):
    from strugatzki_b200 import engine, Aborted
    files, norm = make_db(4, 4000)
    inp = make_input(900)
    op, nc = corr_cfgs(inp, norm, num_matches=3)
    db = build_db(ctx, files, norm)
    job = engine.CorrelationJob(db, nc, inp)
    job.start()
    job.wait()
    p, done, status = job.poll()
    assert done and status == 0 and p == 1.0
    assert_matches_equal(job.result(), O.corr_search(op, files))
    job.abort()
    with pytest.raises(Aborted):
        job.scan()


def test_corr_errors(ctx):
    from strugatzki_b200 import engine, NativeError
    files, norm = make_db(1, 1000)
    inp = make_input(100)
    db = engine.Database(ctx, 14, norm)
    db.add_file(files[0])
    _, nc = corr_cfgs(inp, norm)
    with pytest.raises(NativeError):          # not finalized
        engine.CorrelationJob(db, nc, inp)
    db.finalize()
    with pytest.raises(NativeError) as ei:    # punch span beyond the input file (reference: EOFException)
        engine.CorrelationJob(db, nc, inp)
    assert ei.value.code == N.ERR_IO
    with pytest.raises(NativeError):
        db.add_file(files[0])


@pytest.mark.parametrize("corr_len,weight,num_breaks,min_spacing,span", [
    (22050, 0.5, 20, 22050, (None, None)), (11025, 0.0, 5, 0, (None, None)), (44100, 1.0, 8, 88200, (None, None)),
    (22050, 0.25, 6, 22050, (300 * 512, 2500 * 512)), (22050, 0.5, 3, 22050, (None, 40 * 512)),
    # the break set lives in registers up to 31 breaks (one entry per lane) and in shared memory beyond
    (22050, 0.5, 31, 11025, (None, None)), (22050, 0.5, 32, 11025, (None, None)), (11025, 0.5, 200, 0, (None, None)),
    (22050, 0.5, 1, 22050, (None, None)), (22050, 0.5, 0, 22050, (None, None)),
])
def test_segmentation_bit_identical(ctx, corr_len, weight, num_breaks, min_spacing, span):
    from strugatzki_b200 import engine
    f, cuts = synth.regime_file(synth.BASE_SEED, 5, 4000, 14, 12)
    _, _, _, norm = synth.default_profile(14)
    op = O.SegmParams(step_size=STEP, corr_len=corr_len, temporal_weight=weight, norm=norm, num_breaks=num_breaks,
                      min_spacing=min_spacing, span_start=span[0], span_stop=span[1])
    want, want_curve = O.segm_run(op, f, want_curve=True)
    cfg = N.SegmConfig(STEP, int(span[0] is not None), int(span[1] is not None), span[0] or 0, span[1] or 0,
                       corr_len, weight, num_breaks, min_spacing)
    got, curve, noff = engine.segm_run(ctx, cfg, f, norm, want_curve=True)
    assert np.array_equal(curve.view(np.uint32), want_curve[:noff].view(np.uint32)), "curve is not bit-identical"
    assert [(b["pos"], np.float32(b["sim"]).tobytes()) for b in got] == \
           [(b["pos"], np.float32(b["sim"]).tobytes()) for b in want]


@pytest.mark.parametrize("decim,weight,warp,ceil,inv,cross", [
    (1, 0.5, 1.0, 1.0, False, False), (3, 0.2, 0.5, 0.8, True, False), (2, 1.0, 2.0, 1.0, False, True),
])
def test_selfsimilarity_image_identical(ctx, decim, weight, warp, ceil, inv, cross):
    from strugatzki_b200 import engine
    f1, _ = synth.regime_file(synth.BASE_SEED, 6, 420, 14, 5)
    f2 = synth.regime_file(synth.BASE_SEED, 7, 400, 14, 4)[0] if cross else None
    _, _, _, norm = synth.default_profile(14)
    op = O.SelfParams(step_size=STEP, corr_len=20480, decimation=decim, temporal_weight=weight, norm=norm,
                      color_inv=inv, color_warp=warp, color_ceil=ceil)
    want = O.self_image(op, f1, f2)
    cfg = N.SelfConfig(STEP, 0, 0, 0, 0, 20480, decim, weight, int(inv), warp, ceil, None, 0, 1)   # precise = 1
    got, g = engine.self_run(ctx, cfg, f1, f2, norm)
    assert got.shape == want.shape and g["imgExt"] == want.shape[0]
    assert np.array_equal(got, want)
    # cell-list entry point
    rng = np.random.default_rng(1)
    l = rng.integers(0, g["imgExt"], 64)
    r = rng.integers(0, g["imgExt"], 64)
    sim, rgb = engine.self_cells(ctx, cfg, f1, f2, l, r, norm)
    wsim, wrgb = O.self_cells(op, f1, f2, l, r)
    assert np.array_equal(sim.view(np.uint32), wsim.view(np.uint32)) and np.array_equal(rgb, wrgb)


@pytest.mark.parametrize("frames,corr_len,decim,weight,warp,inv,cross", [
    (900, 44100, 1, 0.5, 1.0, False, False), (1500, 20480, 3, 0.3, 1.0, True, False), (700, 44100, 2, 1.0, 0.5, False, True),
    (650, 10240, 1, 0.0, 1.0, False, False),
    # record grids of the tensor-core kernel: g = gcd(decim, 8), window stride decim / g records (odd strides, strides of
    # 8 frames and more), an H that is a multiple of 16, an H of exactly one K step
    (2300, 44100, 4, 0.5, 1.0, False, False), (3300, 44100, 7, 0.5, 1.0, False, True), (3600, 16384, 8, 0.25, 1.0, False, False),
    (2500, 44100, 5, 0.5, 1.0, True, False), (1900, 44100, 6, 0.7, 1.0, False, False), (700, 8192, 1, 0.5, 1.0, False, False),
    (6000, 44100, 16, 0.5, 1.0, False, False),
    # H = 96: the longest window with one spectral main accumulator (13 x 6 = 78 MMAs); H = 112 and 160: two accumulators
    # by channel parity (7 x 7 and 7 x 10 MMAs), also with the spectral group alone and in expansion mode
    (1500, 49152, 1, 0.5, 1.0, False, False), (1500, 57344, 1, 0.5, 1.0, False, False), (1700, 81920, 1, 0.5, 1.0, False, True),
    (1700, 81920, 2, 0.0, 1.0, False, False), (2400, 65536, 3, 0.5, 1.0, False, False),
    # long windows (32 and 64 K steps per channel): chunks of 160 frames, one launch per chunk; 250 000 samples = H 488 with
    # a last chunk of 8 frames (one K step with a pre-masked tail), in expansion mode too
    (1600, 262144, 1, 0.5, 1.0, False, False), (2600, 524288, 2, 0.4, 1.0, False, True), (1500, 250000, 1, 0.5, 1.0, False, False),
    (2200, 250000, 3, 0.5, 1.0, False, False),
])
def test_selfsimilarity_fast_gram_within_tolerance(ctx, frames, corr_len, decim, weight, warp, inv, cross):
    """default (fast) path: Gram tiles (tensor cores, split FP16) + closed form with FP64-accumulated window sums: sims
    within 1e-5 relative, grey level within 1 LSB"""
    from strugatzki_b200 import engine
    f1, _ = synth.regime_file(synth.BASE_SEED, 16, frames, 14, 6)
    f2 = synth.regime_file(synth.BASE_SEED, 17, frames - 30, 14, 5)[0] if cross else None
    _, _, _, norm = synth.default_profile(14)
    op = O.SelfParams(step_size=STEP, corr_len=corr_len, decimation=decim, temporal_weight=weight, norm=norm,
                      color_inv=inv, color_warp=warp)
    want = O.self_image(op, f1, f2)
    cfg = N.SelfConfig(STEP, 0, 0, 0, 0, corr_len, decim, weight, int(inv), warp, 1.0, None, 0, 0)   # precise = 0
    got, g = engine.self_run(ctx, cfg, f1, f2, norm)
    # no silent fall-back to the FFMA2 kernel: windows whose accumulation chains would exceed the error budget of the
    # truncating tensor-core adder (more than 78 MMAs into one accumulator) run in chunks of 160 frames
    assert engine.self_last_kernel(ctx) == "tc_gram"
    assert got.shape == want.shape and g["imgExt"] == want.shape[0] > 128      # several 128 x 128 tiles
    assert np.array_equal(got, got[::-1, ::-1].T)                              # mirrored like the reference
    dg = np.abs((got & 0xFF).astype(np.int64) - (want & 0xFF).astype(np.int64))
    assert dg.max() <= 1, f"grey level differs by {dg.max()}"
    assert (dg > 0).mean() < 0.01                                              # rounding-boundary flips only
    rng = np.random.default_rng(3)
    l = rng.integers(0, g["imgExt"], 4000)
    r = rng.integers(0, g["imgExt"], 4000)
    sim, _ = engine.self_cells(ctx, cfg, f1, f2, l, r, norm)
    wsim, _ = O.self_cells(op, f1, f2, l, r)
    # relative 1e-5; the absolute floor for sims near zero is 4e-6 (= 0.001 grey levels): the tensor core's truncating
    # accumulation leaves up to 3.4e-6 at H = 86 (tools/selfsim_error_probe.py), the FFMA2 kernel stays below 1e-6
    assert_sims_close(sim, wsim, rel=1e-5, abs_tol=4e-6 if engine.self_last_kernel(ctx) == "tc_gram" else 2e-6,
                      what="selfsim cell")


def test_measured_peaks_are_plausible(ctx):
    ffma = ctx.measure_peak(0)
    hbm = ctx.measure_peak(3)
    assert 30.0 < ffma < 100.0, ffma       # B200: 148 SMs x 128 lanes x 2 flop x ~1.9 GHz ~ 72 TFLOP/s
    assert 3000.0 < hbm < 9000.0, hbm


def test_corr_edge_cases(ctx):
    """empty DB, files shorter than the window, numMatches / numPerFile = 0, and an 8 s window (smaller CTA config)"""
    from strugatzki_b200 import engine
    files, norm = make_db(3, [900, 100, 60])
    inp = make_input(900)
    op, nc = corr_cfgs(inp, norm, num_matches=0)
    assert engine.CorrelationJob(build_db(ctx, files, norm), nc, inp).run() == [] == O.corr_search(op, files)
    # numPerFile = 0 is degenerate in the reference (collapse branch without size check leaks matches): rejected
    _, nc0 = corr_cfgs(inp, norm, num_matches=3, num_per_file=0)
    with pytest.raises(engine.N.NativeError) as ei:
        engine.CorrelationJob(build_db(ctx, files, norm), nc0, inp)
    assert ei.value.code == N.ERR_INVALID
    # handles may be destroyed in any order (garbage-collected hosts): context first, then database, then job
    from strugatzki_b200 import engine as E
    c2 = E.Context(0)
    db2 = E.Database(c2, 14, norm)
    db2.add_file(files[0])
    db2.finalize()
    _, nc3 = corr_cfgs(inp, norm, num_matches=3)
    j2 = E.CorrelationJob(db2, nc3, inp)
    c2.close(); db2.close()
    assert len(j2.run()) == 1          # still fully usable: the library keeps parents alive for their children
    j2.close()
    op, nc = corr_cfgs(inp, norm, num_matches=3)
    assert engine.CorrelationJob(build_db(ctx, [], norm), nc, inp).run() == []
    short = [f[:150] for f in files]                                      # every file shorter than W = 172
    assert engine.CorrelationJob(build_db(ctx, short, norm), nc, inp).run() == [] == O.corr_search(op, short)
    # 8 s punch window: W = 689 frames
    files, norm = make_db(2, [2500, 1800])
    plant_needles(files, inp[:689], [(1, 300)])
    op, nc = corr_cfgs(inp, norm, punch_in=(0, 352800), num_matches=2)
    job = engine.CorrelationJob(build_db(ctx, files, norm), nc, inp)
    got = job.run()
    assert_matches_equal(got, O.corr_search(op, files))
    assert (got[0]["file"], got[0]["start"]) == (1, 300 * STEP)
    want, _ = O.corr_curve(op, files[0])
    sim, _ = job.curve(0, 0, 0, len(want))
    assert_sims_close(sim, want, what="W=689 curve")


@pytest.mark.parametrize("W,weight", [(2, 0.5), (5, 0.5), (16, 0.3), (31, 1.0), (40, 0.0)])
def test_corr_short_windows_and_level_steps(ctx, W, weight):
    """K1 slides its window statistics in FP32 centred on the mean of a thread's first window (corr_tc2.cuh, T2Win).  With
    a window of a few frames the mean moves by a whole window within the 15 slides of a run, and a file whose level jumps
    (quiet -> loud, a factor of 40) moves it by many standard deviations: such windows must be recognised (c1^2 against
    n c2) and re-evaluated exactly, so that the curve still meets the contract everywhere."""
    from strugatzki_b200 import engine
    files, norm = make_db(3, [1500, 700, 2100])
    inp = make_input(300)
    f = files[2]
    f[500:900] = (f[500:900] * np.float32(0.025)).astype(np.float32)       # a quiet stretch with hard edges
    f[1200:1210] = f[1200]                                                 # ten constant frames: exactly constant windows for W <= 10
    plant_needles(files, inp[:W], [(0, 333), (2, 1700)])
    op, nc = corr_cfgs(inp, norm, punch_in=(0, W * STEP), w_in=weight, num_matches=6, num_per_file=2, max_boost=1e9)
    job = engine.CorrelationJob(build_db(ctx, files, norm), nc, inp)
    got = job.run()
    assert_matches_equal(got, O.corr_search(op, files))
    for i, fl in enumerate(files):
        want, _ = O.corr_curve(op, fl)
        sim, _ = job.curve(i, 0, 0, len(want))
        assert_sims_close(sim, want, rel=1e-5, abs_tol=2e-6, what=f"W={W} file {i}")


@pytest.mark.parametrize("W", [258, 1723, 3000])
def test_corr_long_windows_run_in_passes_on_the_tensor_cores(ctx, W):
    """The reference accepts any punch span (FeatureCorrelationImpl.scala:83-98,154).  Beyond 257 frames the tensor-core
    K1 cuts the window into passes of 16 K steps (corr_tc2.cuh, T2Geom): 258 is the first two-pass window, 1723 = 20 s and
    3000 = 35 s do not fit the FFMA2 kernel's shared-memory tile at all (round 1 returned SGZ_ERR_INVALID there)."""
    from strugatzki_b200 import engine
    files, norm = make_db(3, [W + 1500, W + 40, 2 * W + 700])
    inp = make_input(W + 300)
    plant_needles(files, inp[:W], [(0, 1000), (2, W + 77)])
    op, nc = corr_cfgs(inp, norm, punch_in=(0, W * STEP), num_matches=4, num_per_file=2)
    job = engine.CorrelationJob(build_db(ctx, files, norm), nc, inp)
    got = job.run()
    assert_matches_equal(got, O.corr_search(op, files))
    assert {(m["file"], m["start"]) for m in got[:2]} == {(0, 1000 * STEP), (2, (W + 77) * STEP)}
    for i, f in enumerate(files):
        want, _ = O.corr_curve(op, f)
        sim, _ = job.curve(i, 0, 0, len(want))
        assert_sims_close(sim, want, rel=1e-5, abs_tol=2e-6, what=f"W={W} file {i}")


def test_corr_digital_silence_gives_nan_like_the_reference(ctx):
    """a constant loudness stretch has zero variance: the reference returns NaN sims (0/0) and, while the result
    list still has space, inserts them at the head; the engine must do the same"""
    from strugatzki_b200 import engine
    files, norm = make_db(4, 2200)
    inp = make_input(900)
    files[1][0:500, 0] = np.float32(0.4)                                  # > W frames of constant loudness at the file start
    plant_needles(files, inp[:172], [(2, 900)])
    op, nc = corr_cfgs(inp, norm, num_matches=6, num_per_file=2)
    job = engine.CorrelationJob(build_db(ctx, files, norm), nc, inp)
    got = job.run()
    want = O.corr_search(op, files)
    wc, _ = O.corr_curve(op, files[1])
    gc, _ = job.curve(1, 0, 0, len(wc))
    assert np.isnan(wc).sum() > 300 and np.array_equal(np.isnan(wc), np.isnan(gc))
    assert_matches_equal(got, want)
    assert any(np.isnan(m["sim"]) for m in want)


@pytest.mark.parametrize("weight", [0.5, 0.0, 1.0])
def test_corr_near_constant_windows_are_the_references(ctx, weight):
    """the reference returns NaN only for an EXACTLY constant window (0 / 0, MathUtil.scala:195); a window that is constant
    but for one cell one ulp off, or that carries a dither of 1e-5 of its level, gets a finite sim from its two-pass Double
    arithmetic.  The tensor-core K1 cannot resolve such windows (variance below 1e-4 of the mean square): it flags them and
    corr_fix.cuh replays the reference's arithmetic, so NaN pattern AND values are the oracle's, bit for bit"""
    from strugatzki_b200 import engine
    files, norm = make_db(3, 4000)
    inp = make_input(900)
    f = files[1]
    f[600:1500] = f[600]                                                   # 900 constant frames, all channels
    f[1000, 0] = np.nextafter(f[1000, 0], np.float32(2))                   # loudness one ulp off at one frame
    f[1100, 5] = np.nextafter(f[1100, 5], np.float32(2))                   # one spectral cell one ulp off
    rng = np.random.default_rng(5)
    f[2000:2800] = f[2000] * (1 + 1e-5 * rng.standard_normal((800, 14))).astype(np.float32)   # dither
    files[2][3000:] = files[2][3000]                                       # constant up to the file end
    plant_needles(files, inp[:172], [(0, 900), (2, 100)])
    op, nc = corr_cfgs(inp, norm, w_in=weight, num_matches=8, num_per_file=3, min_spacing=22050)
    job = engine.CorrelationJob(build_db(ctx, files, norm), nc, inp)
    got = job.run()
    flagged = 0
    for i in range(3):
        wc, wb = O.corr_curve(op, files[i])
        gc, gb = job.curve(i, 0, 0, len(wc))
        assert np.array_equal(np.isnan(wc), np.isnan(gc)), f"file {i}: NaN pattern"
        assert_sims_close(gc, wc, what=f"file {i} sim")
        if i == 1 and weight > 0:
            # windows inside the flat stretches have an (almost) constant LOUDNESS channel: flagged and replayed exactly.
            # (The spectral group is not ill-conditioned there: its 13 channels sit at different levels.)
            flat = np.zeros(len(wc), bool)
            flat[600:1500 - 171] = True
            flat[2000:2800 - 171] = True
            fin = flat & np.isfinite(wc)          # (NaN patterns were compared above; NaN payloads differ by platform)
            assert np.array_equal(gc[fin].view(np.uint32), wc[fin].view(np.uint32))        # exact replay: bit identical
            assert np.isfinite(wc[flat]).any() and np.isnan(wc[flat]).any()
            flagged = int(flat.sum())
    assert flagged > 1000 or weight == 0
    assert_matches_equal(got, O.corr_search(op, files))


def test_memory_pool_reuses_and_trims(ctx):
    """destroyed databases park their buffers; a new database of the same size gets them back dirty and must still
    produce the same result (slack zeroing at finalize); trim returns the memory to the driver"""
    from strugatzki_b200 import engine
    files, norm = make_db(3, 2500)
    inp = make_input(900)
    op, nc = corr_cfgs(inp, norm, num_matches=4, num_per_file=2)
    want = O.corr_search(op, files)
    ctx.trim()
    for _ in range(3):
        db = build_db(ctx, files, norm)
        job = engine.CorrelationJob(db, nc, inp)
        assert_matches_equal(job.run(), want)
        job.close()
        db.close()
    assert ctx.trim() > 0
    assert ctx.trim() == 0


@pytest.mark.parametrize("punch_out", [False, True])
def test_streaming_scan_behind_async_upload(ctx, punch_out):
    """sgz_db_finalize_async: the search starts while HOST_STABLE uploads are still in flight and K1 runs range by
    range behind the upload markers -- same matches as the resident database and the oracle"""
    import torch
    from strugatzki_b200 import engine
    lens = [3000, 2500, 4100, 180, 2900, 3300, 2000, 5000, 172, 2600, 3100, 2800] * 3
    files, norm = make_db(len(lens), lens)
    inp = make_input(900)
    plant_needles(files, inp[:172], [(2, 1500), (7, 10), (19, 3000), (33, 2400)])
    kw = dict(num_matches=9, num_per_file=2, min_spacing=22050)
    if punch_out:
        kw.update(punch_in=(0, 44100), punch_out=(200 * 512, 200 * 512 + 44100), min_punch=22050, max_punch=88200)
    op, nc = corr_cfgs(inp, norm, **kw)
    want = O.corr_search(op, files)
    pinned = [torch.from_numpy(f).pin_memory() for f in files]     # must outlive the search (HOST_STABLE)
    for rep in range(2):                                           # second round reuses pooled (dirty) buffers
        db = engine.Database(ctx, 14, norm)
        for t in pinned:
            db.add_file_ptr(t.data_ptr(), t.shape[0], N.LAYOUT_INTERLEAVED_LE | N.LAYOUT_HOST_STABLE)
        db.finalize(wait=False)
        job = engine.CorrelationJob(db, nc, inp)
        got = job.run()
        assert_matches_equal(got, want)
        assert job.num_offsets == O.corr_num_offsets(op, lens)
        # the database is fully resident afterwards: a second (monolithic) scan -- on the tensor-core kernel, the
        # streaming one used the FFMA2 kernel -- gives the same curves within the kernels' tolerance
        sim_a, _ = job.curve(2, 0, 0, 1000)
        job.scan()
        sim_b, _ = job.curve(2, 0, 0, 1000)
        assert np.array_equal(np.isnan(sim_a), np.isnan(sim_b)) and np.nanmax(np.abs(sim_a - sim_b)) < 4e-6
        db.finalize()                                              # explicit wait is a no-op now
        job.close()
        db.close()


@pytest.mark.parametrize("n1,n2,w,use_norm,span1,span2", [
    (300, 9000, 0.5, True, (None, None), (None, None)),
    (9300, 172, 0.3, True, (None, None), (None, None)),          # template is the SECOND file
    (300, 8192, 0.5, False, (None, None), (None, None)),         # first read swallows everything: one value
    (2000, 12000, 0.0, True, (100 * 512, 350 * 512), (None, 11000 * 512)),
    (400, 9500, 1.0, True, (None, None), (700 * 512, None)),
    (5000, 8700, 0.5, True, (None, None), (None, None)),         # L > 4096: the read index wraps at 8192
])
def test_cross_similarity_bit_identical(ctx, n1, n2, w, use_norm, span1, span2):
    """CrossSimilarity curve: the engine replays the reference's ring buffer exactly"""
    from strugatzki_b200 import engine
    files, norm = make_db(2, [n1, n2])
    norm = norm if use_norm else None
    want = O.cross_run(O.CrossParams(step_size=STEP, norm=norm, temporal_weight=w, span1=span1, span2=span2),
                       files[0], files[1])
    cfg = N.CrossConfig(STEP, int(span1[0] is not None), int(span1[1] is not None), int(span2[0] is not None),
                        int(span2[1] is not None), 0, span1[0] or 0, span1[1] or 0, span2[0] or 0, span2[1] or 0, w, 8.0)
    got = engine.cross_run(ctx, cfg, files[0], files[1], norm)
    assert got.shape == want.shape and got.shape[0] >= 1
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_cross_similarity_boost_gate_and_errors(ctx):
    from strugatzki_b200 import engine, NativeError
    files, norm = make_db(2, [300, 9000])
    quiet = files[1].copy()
    quiet[:, 0] *= np.float32(0.02)                   # boost > maxBoost -> sim forced to 0f
    cfg = N.CrossConfig(STEP, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.5, 2.0)
    got = engine.cross_run(ctx, cfg, files[0], quiet, None)
    want = O.cross_run(O.CrossParams(step_size=STEP, norm=None, max_boost=2.0), files[0], quiet)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and np.all(got == 0.0)
    big, _ = make_db(2, [9000, 9100])
    with pytest.raises(NativeError) as ei:            # reference: ArrayIndexOutOfBoundsException
        engine.cross_run(ctx, cfg, big[0], big[1], None)
    assert ei.value.code == N.ERR_INVALID


def test_feature_stats_matches_oracle(ctx):
    """FeatureStats: per-file percentiles and the database-wide (p01 min, p99 max) spans.
    Counts are integers and the Double sum keeps the reference's order, so only libm-vs-CUDA pow/log ulps differ."""
    from strugatzki_b200 import engine
    files, _ = make_db(6, [3000, 52, 9000, 20000, 4100, 777])
    files[2][:, 3] = np.float32(0.25)                    # constant channel: d == 0 -> NaN like the reference
    files[3][:, 5] *= np.float32(-3.0)                   # negative values
    want, want_per = O.stats_run(files, want_per_file=True)
    db = engine.Database(ctx, 14, None)
    for i, f in enumerate(files):
        db.add_file(f.astype(">f4"), N.LAYOUT_INTERLEAVED_BE) if i % 2 else db.add_file(f)
    db.finalize()
    got, got_per = db.stats(want_per_file=True)
    assert np.array_equal(np.isnan(got_per), np.isnan(want_per))
    assert np.allclose(got_per, want_per, rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(got, want, rtol=1e-9, atol=0, equal_nan=True)
    assert np.isnan(got[3]).all() and not np.isnan(got[[0, 1, 2, 4]]).any()
    # what ends up in feat_norms.aif is the Float narrowing: identical bits
    ok = ~np.isnan(want)
    assert np.array_equal(got.astype(np.float32)[ok].view(np.uint32), want.astype(np.float32)[ok].view(np.uint32))
    # a one-frame file has d == 0 in every channel: its NaNs poison all spans through math.min / math.max
    tiny, _ = make_db(3, [500, 1, 600])
    dbt = engine.Database(ctx, 14, None)
    for f in tiny:
        dbt.add_file(f)
    dbt.finalize()
    assert np.isnan(dbt.stats()).all() and np.isnan(O.stats_run(tiny)).all()
    from strugatzki_b200 import NativeError
    _, norm = make_db(1, 10)
    dbn = engine.Database(ctx, 14, norm)
    dbn.add_file(files[0])
    dbn.finalize()
    with pytest.raises(NativeError):                     # normalised database: stats need the raw values
        dbn.stats()


@pytest.mark.parametrize("w_in,weight", [(88200, 0.5), (22050, 0.3), (100000, 1.0), (131072, 0.5)])   # last: W = 256, the widest TC window
def test_corr_both_k1_kernels_match_oracle(ctx, w_in, weight, monkeypatch):
    """K1 exists twice: on the tensor cores (default where it applies: tcgen05 split-FP16 MMAs on a Hankel view of the
    channel rows, corr_tc.cuh) and as the FFMA2 kernel (SGZ_CORR_TC=0; wide windows, > 14 channels, streaming scans).
    Same tolerance for both, and they agree with each other."""
    from strugatzki_b200 import engine
    files, norm = make_db(4, [9000, 400, 5100, 7000])
    inp = make_input(900)
    W = (w_in + STEP // 2) // STEP
    plant_needles(files, inp[:W], [(0, 4090), (2, 37), (3, 6000)])
    op, nc = corr_cfgs(inp, norm, punch_in=(0, w_in), w_in=weight, num_matches=5, num_per_file=2)
    db = build_db(ctx, files, norm)
    monkeypatch.setenv("SGZ_CORR_TC", "0")
    ffma_job = engine.CorrelationJob(db, nc, inp)
    monkeypatch.setenv("SGZ_CORR_TC", "1")
    tc_job = engine.CorrelationJob(db, nc, inp)
    monkeypatch.delenv("SGZ_CORR_TC")
    want = O.corr_search(op, files)
    assert_matches_equal(ffma_job.run(), want)
    assert_matches_equal(tc_job.run(), want)
    differs = False
    for i, f in enumerate(files):
        want_sim, want_boost = O.corr_curve(op, f, 0, 0)
        n = len(want_sim)
        if n == 0:
            continue
        sims = []
        for job, name in ((tc_job, "tensor cores"), (ffma_job, "FFMA2")):
            sim, boost = job.curve(i, 0, 0, n)
            assert_sims_close(sim, want_sim, rel=1e-5, abs_tol=2e-6, what=f"file {i} sim ({name})")
            assert_sims_close(boost, want_boost, rel=1e-5, abs_tol=0, what=f"file {i} boost ({name})")
            sims.append(sim)
        assert np.nanmax(np.abs(sims[0] - sims[1])) < 4e-6
        differs |= not np.array_equal(sims[0].view(np.uint32), sims[1].view(np.uint32))
    assert differs                                       # i.e. two different kernels really ran


def test_selfsimilarity_small_image_and_nan_fallback(ctx):
    """An image smaller than one 128 x 128 tile on the tensor-core kernel; non-finite features go to the exact replay
    (the Gram forms centre by the file mean, which a NaN would smear over the whole image) and give the oracle's image: a
    NaN frame only touches the windows over it."""
    from strugatzki_b200 import engine
    _, _, _, norm = synth.default_profile(14)
    f1, _ = synth.regime_file(synth.BASE_SEED, 41, 260, 14, 3)
    op = O.SelfParams(step_size=STEP, corr_len=44100, decimation=1, temporal_weight=0.5, norm=norm)
    cfg = N.SelfConfig(STEP, 0, 0, 0, 0, 44100, 1, 0.5, 0, 1.0, 1.0, None, 0, 0)
    want = O.self_image(op, f1, None)
    got, g = engine.self_run(ctx, cfg, f1, None, norm)
    assert engine.self_last_kernel(ctx) == "tc_gram" and 0 < g["imgExt"] < 128
    d = np.abs((got & 0xFF).astype(np.int64) - (want & 0xFF).astype(np.int64))
    assert got.shape == want.shape and d.max() <= 1
    f2 = f1.copy()
    f2[100, 3] = np.nan
    want = O.self_image(op, f2, None)
    got, _ = engine.self_run(ctx, cfg, f2, None, norm)
    assert engine.self_last_kernel(ctx) == "fp64_replay"
    assert np.array_equal(got, want) and (want == 0).any() and (want != 0).any()
    # no window at all (file shorter than two half windows): an empty image, no launch
    short = f1[:100]
    got, g = engine.self_run(ctx, cfg, short, None, norm)
    assert g["imgExt"] == 0


@pytest.mark.parametrize("num_breaks", [4, 40])
def test_segmentation_with_silence_and_ties(ctx, num_breaks):
    """digital silence gives NaN sims (0/0 in correlateHalf) and repeated material gives exactly equal sims: the replay of
    the reference's TreeSet (add does not overwrite, NaN greatest) must survive both, in both replay kernels"""
    from strugatzki_b200 import engine
    f, _ = synth.regime_file(synth.BASE_SEED, 8, 1500, 14, 5)
    f[300:520] = f[300]                 # constant stretch: windows inside it have zero variance
    f[900:1100] = f[600:800]            # repeated material: equal sims at equal relative positions
    f[1100:1300] = f[600:800]
    _, _, _, norm = synth.default_profile(14)
    op = O.SegmParams(step_size=STEP, corr_len=22050, temporal_weight=0.5, norm=norm, num_breaks=num_breaks, min_spacing=22050)
    want, want_curve = O.segm_run(op, f, want_curve=True)
    cfg = N.SegmConfig(STEP, 0, 0, 0, 0, 22050, 0.5, num_breaks, 22050)
    got, curve, noff = engine.segm_run(ctx, cfg, f, norm, want_curve=True)
    wc = want_curve[:noff]
    assert np.isnan(wc).any()
    assert np.array_equal(np.isnan(curve), np.isnan(wc))               # NaN where the reference has NaN (any payload)
    ok = ~np.isnan(wc)
    assert np.array_equal(curve[ok].view(np.uint32), wc[ok].view(np.uint32)), "curve is not bit-identical"
    # a NaN sim can be a break (Float.compare sorts NaN last, so it survives while the set is not full): same position,
    # NaN on both sides, any payload; every other sim bit for bit
    def key(b):
        x = np.float32(b["sim"])
        return (b["pos"], b"nan" if np.isnan(x) else x.tobytes())
    assert [key(b) for b in got] == [key(b) for b in want]
