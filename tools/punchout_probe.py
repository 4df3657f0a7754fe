"""FeatureCorrelation punch-in + punch-out timing (BASELINE.json configs[2]) on the GPU (developer tool):
python tools/punchout_probe.py [files]; under `ncu --metrics gpu__time_duration.sum` the launch list gives the kernel split."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

STEP, FR = 512, 51680
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 60
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mu, sigma, floor0, norm = synth.default_profile(14)
inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
ctx = engine.Context(0)
db = engine.Database(ctx, 14, norm)
db.reserve(nf * FR, nf)
for i in range(nf):
    db.add_synth(synth.BASE_SEED, 1 + i, FR, mu, sigma, float(floor0))
for k, (f, a, b) in enumerate([(2, 5000, 5400), (33, 30000, 30650)]):
    if f < nf:
        db.patch(f, a, synth.plant(inp[:172], 41, 2 * k))
        db.patch(f, b, synth.plant(inp[345:517], 41, 2 * k + 1))
db.finalize()
cfg = N.CorrConfig(STEP, 0, 88200, 0.5, 1, 176640, 264704, 0.5, 44100, 352800, 8.0, 20, 2, 22050)
job = engine.CorrelationJob(db, cfg, inp)
job.run()
times = []
for _ in range(reps):
    ctx.synchronize(); t = time.perf_counter(); got = job.run(); ctx.synchronize(); times.append(time.perf_counter() - t)
tm = job.timing()
print(json.dumps(dict(files=nf, offsets=job.num_offsets, wall_ms=round(float(np.median(times)) * 1e3, 3),
                      scan_ms=round(tm["scan_ms"], 3), select_ms=round(tm["select_ms"], 3), matches=len(got),
                      last_sim=got[-1]["sim"] if got else None)), flush=True)
