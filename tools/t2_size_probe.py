"""Does the cost per offset of k_corr_tc2 depend on the size of the database?  Scans databases of several sizes (files of the bench's shape) and prints ms per scan and ns per 8192-offset tile per SM; with SGZ_CORR_TC_PROF=1 the role counters per size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from strugatzki_b200 import _native as N, engine, synth
F = bench.FRAMES_PER_FILE
sizes = [int(s) for s in os.environ.get("T2_SIZES", "600,2000,6000").split(",")]
reps = int(os.environ.get("T2_REPS", "20"))
ctx = engine.Context(0)
mu, sigma, floor0, norm = synth.default_profile(14)
inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
for files in sizes:
    db = engine.Database(ctx, 14, norm)
    db.reserve(files * F, files)
    db.add_synth_many(synth.BASE_SEED, 1, files, F, mu, sigma, float(floor0))
    db.finalize()
    job = engine.CorrelationJob(db, bench.corr_config(N), inp)
    for _ in range(int(os.environ.get("T2_WARM", max(reps, 430000 // files)))):      # about half a second of load: clocks ramp up (and the power cap sets in) under continuous load only
        job.scan()
    ms, tail = [], []
    for _ in range(reps):
        job.scan(); ms.append(job.timing()["scan_ms"]); tail.append(job.timing()["select_ms"])
    m = float(np.median(ms))
    tiles = files * F / 8192.0
    print("files", files, "scan_ms", round(m, 4), "min", round(float(np.min(ms)), 4), "us_per_tile_per_sm", round(m * 1e3 / (tiles / 148.0), 3),
          "offsets_per_s", "%.4g" % (files * F / (m * 1e-3)), "tail_ms", round(float(np.median(tail)), 4), flush=True)
    del job, db
