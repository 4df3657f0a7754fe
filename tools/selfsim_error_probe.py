"""Worst deviation of the default SelfSimilarity kernel's sims from the oracle as a function of the window length
(developer tool): python tools/selfsim_error_probe.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

ctx = engine.Context(0)
_, _, _, norm = synth.default_profile(14)
for corr_len, frames in ((22050, 1200), (44100, 1200), (49152, 1400), (57344, 1500), (81920, 1700), (90112, 1800), (98304, 1500)):
    f1, _ = synth.regime_file(synth.BASE_SEED, 16, frames, 14, 6)
    op = O.SelfParams(step_size=512, corr_len=corr_len, decimation=1, temporal_weight=0.5, norm=norm)
    cfg = N.SelfConfig(512, 0, 0, 0, 0, corr_len, 1, 0.5, 0, 1.0, 1.0, None, 0, 0)
    g = engine.self_geometry(cfg, frames, frames)
    rng = np.random.default_rng(3)
    l = rng.integers(0, g["imgExt"], 6000)
    r = rng.integers(0, g["imgExt"], 6000)
    sim, _ = engine.self_cells(ctx, cfg, f1, None, l, r, norm)
    engine.self_run(ctx, cfg, f1, None, norm, download=False)
    kern = engine.self_last_kernel(ctx)
    want, _ = O.self_cells(op, f1, None, l, r)
    err = np.abs(sim.astype(np.float64) - want.astype(np.float64))
    rel = err / np.maximum(np.abs(want), 1e-30)
    print(json.dumps(dict(corr_len=corr_len, H=(corr_len + 256) // 512, kernel=kern, max_abs=float(err.max()),
                          mean_signed=float(np.mean(sim.astype(np.float64) - want)), max_rel_where_abs_gt_2e6=float(rel[err > 2e-6].max()) if (err > 2e-6).any() else 0.0)), flush=True)
