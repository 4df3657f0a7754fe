"""CrossSimilarity timing on the GPU (developer tool): 2 s template over a 1 h feature file."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

mu, sigma, floor0, norm = synth.default_profile(14)
n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 310078
cf = [synth.synth_file(synth.BASE_SEED, 900, 172, mu, sigma, floor0), synth.synth_file(synth.BASE_SEED, 901, n2, mu, sigma, floor0)]
ctx = engine.Context(0)
ccfg = N.CrossConfig(512, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.5, 8.0)
for rep in range(3):
    gc = engine.cross_run(ctx, ccfg, cf[0], cf[1], norm)
    kms, _ = ctx.last_timing()
    print(json.dumps(dict(outputs=int(gc.shape[0]), kernel_ms=round(kms, 3), outputs_per_s=round(gc.shape[0] / (kms * 1e-3), 1))),
          flush=True)
