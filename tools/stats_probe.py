"""FeatureStats timing on the GPU (developer tool): python tools/stats_probe.py [files]

Under `ncu --metrics gpu__time_duration.sum` the launch list gives the split between the three passes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import engine, synth  # noqa: E402

nst = int(sys.argv[1]) if len(sys.argv) > 1 else 600
mu, sigma, floor0, _ = synth.default_profile(14)
ctx = engine.Context(0)
raw = engine.Database(ctx, 14, None)
raw.reserve(nst * 51680, nst)
for i in range(nst):
    raw.add_synth(synth.BASE_SEED, 5000 + i, 51680, mu, sigma, float(floor0))
raw.finalize()
for rep in range(3):
    raw.stats()
    ms, launches = ctx.last_timing()
    frames = nst * 51680
    print(json.dumps(dict(files=nst, frames=frames, kernels_ms=round(ms, 3), launches=int(launches),
                          frames_per_s=round(frames / (ms * 1e-3), 1),
                          algorithmic_GBps=round(frames * 112 / (ms * 1e-3) / 1e9, 1))), flush=True)
