"""SelfSimilarity cells/s on the GPU (developer tool): python tools/selfsim_probe.py [frames] [decim]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
decim = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = engine.Context(0)
f, _ = synth.regime_file(synth.BASE_SEED, 4, frames, 14, max(4, frames // 2000))
_, _, _, norm = synth.default_profile(14)
cfg = N.SelfConfig(512, 0, 0, 0, 0, 44100, decim, 0.5, 0, 1.0, 1.0, None, 0)
for rep in range(2):
    t = time.time()
    _, g = engine.self_run(ctx, cfg, f, None, norm, download=False)
    wall = time.time() - t
    ms, launches = ctx.last_timing()
    print(json.dumps(dict(frames=frames, imgExt=g["imgExt"], decim=g["decim"], cells=g["numCells"], kernel_ms=round(ms, 2),
                          wall_ms=round(wall * 1e3, 1), cells_per_s=round(g["numCells"] / (ms * 1e-3), 1), kernel=engine.self_last_kernel(ctx))), flush=True)
