"""Where does the end-to-end (host buffers -> matches) step spend its time?  Phase timing of bench.py's e2e path."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

files = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
STREAM = int(sys.argv[2]) if len(sys.argv) > 2 else 0
F = bench.FRAMES_PER_FILE
ctx = engine.Context(0)
mu, sigma, floor0, norm = synth.default_profile(14)
inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
cfg = bench.corr_config(N)
host = torch.empty((files, 14, F), dtype=torch.float32, pin_memory=True)
host.normal_()
base = host.data_ptr()
bpf = F * 14 * 4
for rep in range(3):
    ph = {}
    ctx.synchronize()
    t0 = t = time.perf_counter()

    def mark(k, sync=True):
        global t
        if sync:
            ctx.synchronize()
        now = time.perf_counter()
        ph[k] = 1e3 * (now - t)
        t = now
    db = engine.Database(ctx, 14, norm)
    db.reserve(files * F, files)
    mark("reserve")
    for i in range(files):
        db.add_file_ptr(base + i * bpf, F, N.LAYOUT_PLANAR_LE | N.LAYOUT_HOST_STABLE)
    mark("add_issue", sync=False)
    if STREAM == 0:
        mark("add_drain")
    db.finalize(wait=STREAM == 0)
    mark("finalize", sync=STREAM == 0)
    job = engine.CorrelationJob(db, cfg, inp)
    mark("job_create", sync=STREAM == 0)
    job.run()
    mark("run")
    job.close()
    db.close()
    mark("close")
    ph["total"] = 1e3 * (time.perf_counter() - t0)
    ph["upload_GBps"] = files * bpf / 1e6 / (ph["add_issue"] + ph.get("add_drain", ph["run"]))
    print(json.dumps(ph))
