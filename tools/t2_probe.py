"""K1 variants on the same resident database: N = 64 tensor-core kernel (default), round-1 N = 32 kernel (SGZ_CORR_TC2=0),
FFMA2 kernel (SGZ_CORR_TC=0): curve agreement against the FFMA2 kernel, scan time.  usage: t2_probe.py [files] [modes]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

files = int(sys.argv[1]) if len(sys.argv) > 1 else 600
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["ffma", "tc1", "tc2"]
F = bench.FRAMES_PER_FILE
ctx = engine.Context(0)
mu, sigma, floor0, norm = synth.default_profile(14)
db = engine.Database(ctx, 14, norm)
db.reserve(files * F, files)
db.add_synth_many(synth.BASE_SEED, 1, files, F, mu, sigma, float(floor0))
inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
db.patch(files // 2, 1234, synth.plant(inp[:172], 77, 0))
db.finalize()
cfg = bench.corr_config(N)
out = {"files": files}
curves = {}
env = {"ffma": {"SGZ_CORR_TC": "0"}, "tc1": {"SGZ_CORR_TC2": "0"}, "tc2": {}}
for mode in modes:
    for k in ("SGZ_CORR_TC", "SGZ_CORR_TC2"):
        os.environ.pop(k, None)
    os.environ.update(env[mode])
    job = engine.CorrelationJob(db, cfg, inp)
    for _ in range(3):
        res = job.run()
    ms = []
    for _ in range(5):
        job.scan()
        ms.append(job.timing()["scan_ms"])
    out["offsets"] = job.num_offsets
    out[mode] = {"scan_ms": float(np.median(ms)), "offsets_per_s": job.num_offsets / (np.median(ms) * 1e-3),
                 "top": res[0], "matches": len(res)}
    curves[mode] = [job.curve(f, 0, 0, F - 171) for f in (0, files // 2, files - 1)]
    job.close()
ref = curves.get("ffma")
if ref is not None:
    for mode in modes:
        if mode == "ffma":
            continue
        worst = 0.0
        for (s0, b0), (s1, b1) in zip(ref, curves[mode]):
            assert np.array_equal(np.isnan(s0), np.isnan(s1)), mode
            worst = max(worst, float(np.nanmax(np.abs(s0 - s1))))
            assert np.allclose(b0, b1, rtol=1e-6, equal_nan=True), mode
        out[mode]["max_abs_sim_diff_vs_ffma"] = worst
print(json.dumps(out))
