"""Shared helpers of the parity tests: synthetic DBs, config twins for oracle and engine."""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402  (tests are allowed to use the oracle)
from strugatzki_b200 import _native as N  # noqa: E402
from strugatzki_b200 import synth  # noqa: E402

STEP = 512


def make_db(num_files: int, n_frames, num_ch: int = 14, seed: int = synth.BASE_SEED, first_stream: int = 1):
    mu, sigma, floor0, norm = synth.default_profile(num_ch)
    lens = [n_frames] * num_files if np.isscalar(n_frames) else list(n_frames)
    files = [synth.synth_file(seed, first_stream + i, lens[i], mu, sigma, floor0) for i in range(num_files)]
    return files, norm


def make_input(n_frames: int = 800, num_ch: int = 14, seed: int = synth.BASE_SEED):
    mu, sigma, floor0, _ = synth.default_profile(num_ch)
    return synth.synth_file(seed, 0, n_frames, mu, sigma, floor0)


def plant_needles(files: List[np.ndarray], window: np.ndarray, where: Sequence[Tuple[int, int]], seed: int = 99,
                  amp: float = 0.02):
    for k, (f, off) in enumerate(where):
        files[f][off:off + window.shape[0]] = synth.plant(window, seed, 1000 + k, amp)


def corr_cfgs(inp, norm, punch_in=(0, 88200), w_in=0.5, punch_out=None, w_out=0.5, min_punch=44100,
              max_punch=352800, max_boost=8.0, num_matches=10, num_per_file=1, min_spacing=0, step=STEP):
    """(oracle CorrParams, native CorrConfig) describing the same FeatureCorrelation.Config."""
    op = O.CorrParams(step_size=step, input=inp, punch_in=punch_in, punch_in_weight=w_in, punch_out=punch_out,
                      punch_out_weight=w_out, min_punch=min_punch, max_punch=max_punch, norm=norm,
                      max_boost=max_boost, num_matches=num_matches, num_per_file=num_per_file,
                      min_spacing=min_spacing)
    po = punch_out or (0, 0)
    nc = N.CorrConfig(step, punch_in[0], punch_in[1], w_in, 0 if punch_out is None else 1, po[0], po[1], w_out,
                      min_punch, max_punch, max_boost, num_matches, num_per_file, min_spacing)
    return op, nc


def build_db(ctx, files, norm):
    from strugatzki_b200 import engine
    db = engine.Database(ctx, files[0].shape[1] if files else 14, norm)
    for f in files:
        db.add_file(f)
    db.finalize()
    return db


def assert_sims_close(got, want, rel=1e-5, abs_tol=2e-6, what="sim") -> int:
    """|got - want| <= max(rel * |want|, abs_tol) everywhere, identical NaN pattern.  Returns how many cells needed the
    ABSOLUTE floor, i.e. missed the relative bound (sims close to zero: the contract's 1e-5 is relative)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs"
    ok = ~nan_w
    err = np.abs(got[ok] - want[ok])
    tol = np.maximum(rel * np.abs(want[ok]), abs_tol)
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} of {ok.sum()} beyond tolerance; worst abs err "
                           f"{err.max():.3e}, worst rel {np.max(err / np.maximum(np.abs(want[ok]), 1e-30)):.3e}")
    return int(np.count_nonzero(err > rel * np.abs(want[ok])))


def assert_matches_equal(got: List[dict], want: List[dict], rel=1e-5, exact_sim=False):
    """same files and spans in the same order; sims and boosts within rel (absolute floor 2e-6).  exact_sim: the sims are
    the oracle's bit for bit (punch-in searches: the decisive offsets are re-evaluated in Double, corr_refine.cuh)"""
    assert len(got) == len(want), f"{len(got)} matches, oracle has {len(want)}\n got={got}\nwant={want}"
    for i, (g, w) in enumerate(zip(got, want)):
        assert (g["file"], g["start"], g["stop"]) == (w["file"], w["start"], w["stop"]), \
            f"match {i}: got {g}, oracle {w}"
        if exact_sim:
            a, b = np.float32(g["sim"]), np.float32(w["sim"])
            assert a.tobytes() == b.tobytes() or (np.isnan(a) and np.isnan(b)), f"match {i} sim: got {a!r}, oracle {b!r}"
        for key in ("sim", "boostIn", "boostOut"):
            a, b = float(g[key]), float(w[key])
            if np.isnan(b):
                assert np.isnan(a)
            else:
                assert abs(a - b) <= max(rel * abs(b), 2e-6), f"match {i} {key}: got {a}, oracle {b}"
