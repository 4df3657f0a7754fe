"""The oracle (oracle/sgz_oracle.c) against an INDEPENDENT numpy formulation (closed forms of SURVEY.md
sections 3.1-3.3), against the Java collection semantics it restates, and against committed golden fixtures.
PARITY UNPINNED by the reference itself (no numeric tests upstream) -- these are the pins we can have."""
import json
import os

import numpy as np
import pytest

from util import O, ROOT, STEP, corr_cfgs, make_db, make_input, plant_needles, synth

GOLDEN = os.path.join(ROOT, "tests", "golden")


def np_corr_curve(inp, f, norm, W, weight, max_boost=8.0):
    """closed form: Pearson of the flattened [C x W] windows, pooled mean/std per group, float blend"""
    def nz(x):
        return ((x - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32) if norm is not None else x
    a = nz(inp[:W]).astype(np.float64)
    b = nz(f).astype(np.float64)
    n = b.shape[0] - W + 1
    sims = np.zeros(n, np.float32)
    boosts = np.zeros(n, np.float32)
    ln_in = np.log(np.float64(np.float32(a[:, 0].mean())))
    for t in range(n):
        w = b[t:t + W]
        boost = np.float32(np.exp((ln_in - np.log(np.float64(np.float32(w[:, 0].mean())))) / 0.6))
        boosts[t] = boost
        if boost <= max_boost:
            def pear(x, y):
                x = x.ravel() - x.mean(); y = y.ravel() - y.mean()
                return np.float32((x * y).sum() / (np.sqrt((x * x).mean()) * np.sqrt((y * y).mean()) * x.size))
            tc = pear(a[:, :1], w[:, :1]) if weight > 0 else np.float32(0)
            sc = pear(a[:, 1:], w[:, 1:]) if weight < 1 else np.float32(0)
            sims[t] = tc * np.float32(weight) + sc * (np.float32(1) - np.float32(weight))
    return sims, boosts


@pytest.mark.parametrize("weight,use_norm", [(0.5, True), (0.0, True), (1.0, False), (0.3, False)])
def test_corr_curve_closed_form(weight, use_norm):
    files, norm = make_db(1, 700)
    inp = make_input(300)
    nrm = norm if use_norm else None
    op, _ = corr_cfgs(inp, nrm, punch_in=(0, 40 * STEP), w_in=weight)
    sim, boost = O.corr_curve(op, files[0])
    wsim, wboost = np_corr_curve(inp, files[0], nrm, 40, weight)
    assert len(sim) == 661
    np.testing.assert_allclose(sim, wsim, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(boost, wboost, rtol=1e-6)


def test_segmentation_and_selfsim_closed_form():
    f, _ = synth.regime_file(synth.BASE_SEED, 3, 600, 14, 5)
    _, _, _, norm = synth.default_profile(14)
    H = 20
    x = ((f - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32).astype(np.float64)

    def half(xl, xr):   # (G - T) / ((Qi + Qj)/2 - T), SURVEY.md 3.3
        n = xl.size
        g = (xl * xr).sum(); s = xl.sum() + xr.sum(); q = (xl ** 2).sum() + (xr ** 2).sum()
        t = s * s / (4 * n)
        return np.float32((g - t) / (q / 2 - t))

    breaks, curve = O.segm_run(O.SegmParams(step_size=STEP, corr_len=H * STEP, norm=norm, num_breaks=4,
                                            min_spacing=0), f, want_curve=True)
    n_off = 600 - 2 * H + 1
    want = np.array([np.float32(0.5) * half(x[t:t + H, :1], x[t + H:t + 2 * H, :1]) +
                     np.float32(0.5) * half(x[t:t + H, 1:], x[t + H:t + 2 * H, 1:]) for t in range(n_off)], np.float32)
    np.testing.assert_allclose(curve[:n_off], want, rtol=3e-6, atol=2e-7)
    assert len(breaks) == 4 and breaks == sorted(breaks, key=lambda b: b["sim"])
    assert breaks[0]["sim"] == curve[:n_off].min()
    assert breaks[0]["pos"] == (int(np.argmin(curve[:n_off])) + H) * STEP

    sp = O.SelfParams(step_size=STEP, corr_len=H * STEP, decimation=3, norm=norm)
    l = np.array([0, 5, 17, 100]); r = np.array([0, 50, 17, 180])
    sim, rgb = O.self_cells(sp, f, None, l, r)
    want = [np.float32(0.5) * half(x[3 * a:3 * a + H, :1], x[3 * b:3 * b + H, :1]) +
            np.float32(0.5) * half(x[3 * a:3 * a + H, 1:], x[3 * b:3 * b + H, 1:]) for a, b in zip(l, r)]
    np.testing.assert_allclose(sim, want, rtol=3e-6, atol=2e-7)
    assert sim[0] == pytest.approx(1.0, abs=1e-6) and rgb[0] == 0xFFFFFF
    img = O.self_image(sp, f)
    assert img.shape == (187, 187) and np.array_equal(img, img[::-1, ::-1].T)   # mirrored about the anti-diagonal
    assert img[186 - 50, 5] == rgb[1] and img[186 - 5, 50] == rgb[1]


def _files_with_sims(sim_lists, W=4):
    """DB whose punch-in curve has chosen peaks: windows equal to the (noisy) query at chosen strengths."""
    rng = np.random.default_rng(0)
    inp = rng.random((W + 2, 3)).astype(np.float32)
    files = []
    for sims in sim_lists:
        f = rng.random((len(sims) * 40 + W, 3)).astype(np.float32) * 0.01
        for k, s in enumerate(sims):
            noise = rng.standard_normal((W, 3)).astype(np.float32)
            f[k * 40:k * 40 + W] = inp[:W] * np.float32(s) + noise * np.float32(1 - s) * 0.2
        files.append(f)
    return inp, files


def test_java_collection_semantics():
    """Q3: once allPrio is full a later, weaker peak of the same file is kept only if it arrives first."""
    W = 4
    inp, files = _files_with_sims([[0.99, 0.98], [0.6, 0.97], [0.97, 0.6]], W)
    p = O.CorrParams(step_size=1, input=inp, punch_in=(0, W), norm=None, max_boost=1e9, num_matches=2,
                     num_per_file=2, min_spacing=0)
    base = O.corr_search(p, files[:1])
    assert [m["file"] for m in base] == [0, 0]
    theta = base[-1]["sim"]
    a = O.corr_search(p, [files[0], files[1]])
    b = O.corr_search(p, [files[0], files[2]])
    # both candidate files hold one peak above theta (else the fixture is broken) -- order decides what survives
    assert all(m["sim"] >= theta for m in a) and all(m["sim"] >= theta for m in b)
    assert len(a) == 2 and len(b) == 2
    # results are sorted by descending sim under Float.compare
    for res in (a, b):
        assert [m["sim"] for m in res] == sorted((m["sim"] for m in res), reverse=True)


def test_spacing_collapse_and_limits():
    files, norm = make_db(3, 1200)
    inp = make_input(400)
    plant_needles(files, inp[:60], [(1, 100), (1, 130), (1, 700)])
    base = dict(punch_in=(0, 60 * STEP), num_matches=6, num_per_file=3)
    op, _ = corr_cfgs(inp, norm, min_spacing=0, **base)
    res = O.corr_search(op, files)
    per_file = {}
    for m in res:
        per_file.setdefault(m["file"], []).append(m)
    assert all(len(v) <= 3 for v in per_file.values()) and len(res) <= 6
    # punch-in spans are W long: two matches of one file never overlap, even with minSpacing = 0
    for v in per_file.values():
        v.sort(key=lambda m: m["start"])
        assert all(b["start"] - a["stop"] >= 0 for a, b in zip(v, v[1:]))
    starts1 = sorted(m["start"] // STEP for m in per_file[1])
    assert 700 in starts1 and (100 in starts1) != (130 in starts1)     # 100 and 130 collapse into one
    op2, _ = corr_cfgs(inp, norm, min_spacing=10 ** 9, **base)
    res2 = O.corr_search(op2, files)
    assert len(res2) == 3 and sorted(m["file"] for m in res2) == [0, 1, 2]   # huge spacing -> one per file


def test_punch_out_geometric_mean_and_range():
    files, norm = make_db(2, 2500)
    inp = make_input(900)
    W = 40
    files[1][500:500 + W] = synth.plant(inp[:W], 1, 1)
    files[1][800:800 + W] = synth.plant(inp[300:300 + W], 1, 2)
    op, _ = corr_cfgs(inp, norm, punch_in=(0, W * STEP), punch_out=(300 * STEP, (300 + W) * STEP),
                      min_punch=100 * STEP, max_punch=600 * STEP, num_matches=2, num_per_file=1)
    res = O.corr_search(op, files)
    assert res[0]["file"] == 1 and res[0]["start"] == 500 * STEP and res[0]["stop"] == 800 * STEP
    sin, _ = O.corr_curve(op, files[1], 0)
    sout, _ = O.corr_curve(op, files[1], 1)
    assert res[0]["sim"] == np.float32(np.sqrt(np.float64(np.float32(sin[500] * sout[800]))))
    assert (res[0]["stop"] - res[0]["start"]) // STEP in range(100, 601)


def test_golden_fixtures():
    """committed outputs of the oracle on seeded inputs (tests/golden/make_golden.py): guards the oracle
    itself against accidental edits."""
    meta = json.load(open(os.path.join(GOLDEN, "golden.json")))
    data = np.load(os.path.join(GOLDEN, "golden.npz"))
    from golden.make_golden import cases
    for name, fn in cases().items():
        got = fn()
        for key, val in got.items():
            want = data[f"{name}.{key}"]
            assert np.array_equal(np.asarray(val).view(np.uint8), want.view(np.uint8)), f"{name}.{key} changed"
    assert sorted(meta["cases"]) == sorted(cases().keys())


def _cross_numpy(f1, f2, norm, w, step=512):
    """independent formulation of CrossSimilarityImpl's ring (see strugatzki_b200/csrc/cross.cuh): simulate the
    8192-frame buffer with plain numpy writes, correlate with float64 numpy reductions"""
    nz = lambda f: ((f - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32) if norm is not None else f
    a, b = (f1, f2) if f1.shape[0] < f2.shape[0] else (f2, f1)
    a, b = nz(a), nz(b)
    L, len2 = a.shape[0], b.shape[0]
    c0 = min(len2, 8192)
    buf = np.zeros((8192, a.shape[1]), np.float32)
    buf[:c0] = b[:c0]
    out = []

    def corr(x, bstat, bwin):
        x = x.astype(np.float64)
        return ((x - x.mean()) * (bwin.astype(np.float64) - bstat.astype(np.float64).mean())).sum() / (
            x.std() * bstat.astype(np.float64).std() * x.size)
    for k in range(1 + len2 - c0):
        if k >= 1:
            buf[(c0 + k - 1) % L] = b[c0 + k - 1]
        idx = (np.arange(L) + k % L) % 8192
        t = np.float32(corr(a[:, :1], buf[:L, :1], buf[idx][:, :1]))
        s = np.float32(corr(a[:, 1:], buf[:L, 1:], buf[idx][:, 1:]))
        out.append(t * np.float32(w) + s * (np.float32(1) - np.float32(w)))
    return np.array(out, np.float32)


@pytest.mark.parametrize("n1,n2,w", [(300, 8700, 0.5), (8600, 172, 0.3), (200, 500, 0.5)])
def test_cross_similarity_ring_semantics(n1, n2, w):
    files, norm = make_db(2, [n1, n2])
    got = O.cross_run(O.CrossParams(step_size=512, norm=norm, temporal_weight=w), files[0], files[1])
    want = _cross_numpy(files[0], files[1], norm, w)
    assert got.shape == want.shape == (1 + max(n1, n2) - min(max(n1, n2), 8192),)
    assert np.allclose(got, want, rtol=2e-6, atol=2e-7)
