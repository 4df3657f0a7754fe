"""Ablation of k_corr_tc2 (profiling build): SGZ_T2_DBG=1 no MMAs, 2 no per-offset epilogue work, 3 both -- what the producer / TMA path, the tensor pipe and the epilogue cost on their own.  Prints the role counters and the scan time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("T2_ABLATE_PROF", "0") == "1":
    os.environ["SGZ_CORR_TC_PROF"] = "1"
import numpy as np
import bench
from strugatzki_b200 import _native as N, engine, synth
files = 600
F = bench.FRAMES_PER_FILE
ctx = engine.Context(0)
mu, sigma, floor0, norm = synth.default_profile(14)
db = engine.Database(ctx, 14, norm)
db.reserve(files * F, files)
db.add_synth_many(synth.BASE_SEED, 1, files, F, mu, sigma, float(floor0))
inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
db.finalize()
cfg = bench.corr_config(N)
job = engine.CorrelationJob(db, cfg, inp)
for _ in range(int(os.environ.get("T2_ABLATE_WARM", "20"))):      # clocks ramp up under continuous load only
    job.scan()
ms = []
for _ in range(int(os.environ.get("T2_ABLATE_REPS", "20"))):
    job.scan(); ms.append(job.timing()["scan_ms"])
print("dbg", os.environ.get("SGZ_T2_DBG"), "scan_ms", float(np.median(ms)), flush=True)
