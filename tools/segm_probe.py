"""FeatureSegmentation timing on the GPU (developer tool): python tools/segm_probe.py [frames]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 51680
ctx = engine.Context(0)
seg, _ = synth.regime_file(synth.BASE_SEED, 31, frames, 14, 26)
_, _, _, norm = synth.default_profile(14)
cfg = N.SegmConfig(512, 0, 0, 0, 0, 22050, 0.5, 20, 22050)
for rep in range(3):
    br, _, noff = engine.segm_run(ctx, cfg, seg, norm, want_curve=True)
    ms, launches = ctx.last_timing()
    print(json.dumps(dict(frames=frames, offsets=int(noff), breaks=len(br), kernel_ms=round(ms, 3), launches=int(launches),
                          offsets_per_s=round(noff / (ms * 1e-3), 1))), flush=True)
