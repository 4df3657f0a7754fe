"""Seeded differential tests: FeatureCorrelation searches with randomly drawn configurations and databases against the
oracle (FeatureCorrelationImpl.scala:73-421).  Every case draws the channel count, ragged file lengths (files shorter than
the punch among them), the punch spans with sample positions that are NOT multiples of the step (the (x + step/2) / step
rounding of :154), weights, minPunch / maxPunch, maxBoost, numMatches / numPerFile / minSpacing, normalisation on or off,
planted needles, repeated files and constant (silent) stretches.  The seeds are fixed, so a case is reproducible by its id."""
import numpy as np
import pytest

from util import N, O, STEP, assert_matches_equal, assert_sims_close, build_db, corr_cfgs, synth

pytestmark = pytest.mark.gpu


def draw_case(seed: int, punch_out: bool):
    rng = np.random.default_rng(1000 + seed)
    num_ch = int(rng.choice([2, 3, 5, 14, 14, 14, 21]))
    mu, sigma, floor0, norm = synth.default_profile(num_ch)
    w_in = int(rng.choice([2, 3, 7, 16, 17, 43, 64, 86, 172, 172, 200, 257, 300]))
    in0 = int(rng.integers(0, 200))
    w_out = int(rng.choice([2, 5, 16, 33, 86, 172, 172, 260])) if punch_out else 0
    gap = int(rng.integers(0, 120))
    out0 = in0 + w_in + gap
    inp = synth.synth_file(synth.BASE_SEED + seed, 0, out0 + w_out + int(rng.integers(1, 60)), mu, sigma, floor0)

    def span(a_frames, n_frames):   # sample positions that round to these frames but are not multiples of the step
        j0, j1 = (int(rng.integers(-STEP // 2 + 1, STEP // 2)) for _ in range(2))
        return (max(a_frames * STEP + j0, 0), (a_frames + n_frames) * STEP + j1)

    n_files = int(rng.integers(1, 12))
    lens = []
    for _ in range(n_files):
        kind = rng.integers(0, 10)
        if kind == 0:
            lens.append(int(rng.integers(1, w_in + 2)))                 # shorter than (or just as long as) the punch
        elif kind == 1:
            lens.append(w_in + w_out + int(rng.integers(0, 40)))
        else:
            lens.append(int(rng.integers(300, 2600)))
    files = [synth.synth_file(synth.BASE_SEED + seed, 1 + i, lens[i], mu, sigma, floor0) for i in range(n_files)]
    # needles: the punch-in window (and, further on, the punch-out window) of the input with a little noise
    for k in range(int(rng.integers(0, 5))):
        f = int(rng.integers(0, n_files))
        length = int(rng.integers(max(w_out, 1), 400)) if punch_out else 0
        if lens[f] < w_in + length + w_out + 2:
            continue
        a = int(rng.integers(0, lens[f] - (w_in + length + w_out) - 1))
        files[f][a:a + w_in] = synth.plant(inp[in0:in0 + w_in], 70 + seed, 2 * k)
        if punch_out:
            b = a + length
            files[f][b:b + w_out] = synth.plant(inp[out0:out0 + w_out], 70 + seed, 2 * k + 1)
    extras = int(rng.integers(0, 4))
    if extras == 1 and n_files > 1 and not punch_out:                   # a file that occurs twice: equal sims, ONE entry of
        files[-1] = files[0].copy()                                     # the reference's TreeSet (punch-in searches
                                                                        # reproduce that: corr_refine.cuh; the cells of a
                                                                        # punch-out search carry the kernel's rounding)
    if extras == 2:                                                     # a constant stretch: NaN windows
        f = int(rng.integers(0, n_files))
        if lens[f] > 40:
            a = int(rng.integers(0, lens[f] - 30))
            files[f][a:a + int(rng.integers(w_in + 1, w_in + 200))] = files[f][a]
    if extras == 3:                                                     # a very quiet file: boost beyond maxBoost
        files[int(rng.integers(0, n_files))][:, 0] *= np.float32(rng.choice([0.02, 0.3]))
    use_norm = bool(rng.integers(0, 4))
    min_punch = int(rng.integers(1, 300))
    cfg = dict(punch_in=span(in0, w_in), w_in=float(rng.choice([0.0, 1.0, 0.5, 0.5, rng.uniform(0.05, 0.95)])),
               punch_out=span(out0, w_out) if punch_out else None,
               w_out=float(rng.choice([0.0, 1.0, 0.5, rng.uniform(0.05, 0.95)])),
               min_punch=min_punch * STEP + int(rng.integers(-200, 200)),
               max_punch=(min_punch + int(rng.integers(0, 700))) * STEP + int(rng.integers(-200, 200)),
               max_boost=float(rng.choice([8.0, 8.0, 2.0, 1.1, 30.0])),
               num_matches=int(rng.choice([1, 2, 3, 5, 10, 30, 100])), num_per_file=int(rng.choice([1, 1, 2, 3, 7])),
               min_spacing=int(rng.choice([0, 0, 512, 22050, 100000, -5])))
    return inp, files, (norm if use_norm else None), cfg


def run_case(ctx, seed: int, punch_out: bool):
    from strugatzki_b200 import engine
    inp, files, norm, cfg = draw_case(seed, punch_out)
    op, nc = corr_cfgs(inp, norm, **cfg)
    want = O.corr_search(op, files)
    db = build_db(ctx, files, norm)
    job = engine.CorrelationJob(db, nc, inp)
    got = job.run()
    assert job.num_offsets == O.corr_num_offsets(op, [f.shape[0] for f in files])
    try:
        assert_matches_equal(got, want, exact_sim=not punch_out)
    except AssertionError:
        # a different list is only acceptable when the curves themselves agree (then it is a near-tie of two sims inside
        # the kernel's rounding, which the caller inspects): report which of the two it was
        for i, f in enumerate(files):
            for which in ((0, 1) if punch_out else (0,)):
                w_sim, _ = O.corr_curve(op, f, which, 0)
                n = len(w_sim) if which or not punch_out else f.shape[0] - nc.minPunch // STEP - 400
                n = min(len(w_sim), max(n, 0))
                if n > 0:
                    g_sim, _ = job.curve(i, which, 0, n)
                    assert_sims_close(g_sim, w_sim[:n], what=f"seed {seed}: file {i} curve {which}")
        raise
    return len(got)


@pytest.mark.parametrize("seed", range(120))
def test_fuzz_punch_in_search(ctx, seed):
    run_case(ctx, seed, punch_out=False)


@pytest.mark.parametrize("seed", range(200, 280))
def test_fuzz_punch_out_search(ctx, seed):
    run_case(ctx, seed, punch_out=True)


@pytest.mark.parametrize("num_matches,num_per_file", [(1, 1), (20, 1), (100, 1), (224, 1), (230, 2), (60, 3)])
def test_many_files_exact_result(ctx, num_matches, num_per_file):
    """more files than matches: the threshold of the exact re-evaluation comes from the numMatches-th largest DISTINCT file
    maximum (k_refine_threshold).  Files that occur two and three
    times take ONE place each in the reference's TreeSet, so the threshold has to reach further down than their copies."""
    from strugatzki_b200 import engine
    rng = np.random.default_rng(5)
    mu, sigma, floor0, norm = synth.default_profile(14)
    n_files, w = 700, 43
    inp = synth.synth_file(synth.BASE_SEED, 0, 120, mu, sigma, floor0)
    files = [synth.synth_file(synth.BASE_SEED, 1 + i, int(rng.integers(150, 420)), mu, sigma, floor0) for i in range(n_files)]
    for k in range(40):                                  # needles of graded quality
        f = int(rng.integers(0, n_files))
        a = int(rng.integers(0, files[f].shape[0] - w))
        files[f][a:a + w] = synth.plant(inp[10:10 + w], 91, k, 0.01 + 0.01 * k)
    for k in range(30):                                  # repeated files, some of them among the best
        src, dst = int(rng.integers(0, n_files)), int(rng.integers(0, n_files))
        files[dst] = files[src].copy()
    op, nc = corr_cfgs(inp, norm, punch_in=(10 * STEP, (10 + w) * STEP), num_matches=num_matches, num_per_file=num_per_file,
                       min_spacing=4 * STEP)
    want = O.corr_search(op, files)
    got = engine.CorrelationJob(build_db(ctx, files, norm), nc, inp).run()
    assert_matches_equal(got, want, exact_sim=True)
    assert len(got) == min(num_matches, len(want)) and len(got) >= min(num_matches, 200)
