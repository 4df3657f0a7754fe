"""Quick on-GPU probe: pipe peaks + K1 throughput on a mid-size synthetic DB (developer tool)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402


def main():
    files = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 51680
    ctx = engine.Context(0)
    out = {}
    for name, which in (("ffma_tflops", 0), ("ffma2_tflops", 1), ("dfma_tflops", 2), ("hbm_copy_gbs", 3),
                        ("lds128_gbs", 4)):
        out[name] = round(ctx.measure_peak(which), 2)
    print(json.dumps(out), flush=True)
    mu, sigma, floor0, norm = synth.default_profile(14)
    db = engine.Database(ctx, 14, norm)
    db.reserve(files * frames, files)
    t = time.time()
    for i in range(files):
        db.add_synth(synth.BASE_SEED, 1 + i, frames, mu, sigma, float(floor0))
    db.finalize()
    print("synth+finalize s", round(time.time() - t, 3), db.info(), flush=True)
    inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
    cfg = N.CorrConfig(512, 0, 88200, 0.5, 0, 0, 0, 0.5, 44100, 352800, 8.0, 100, 1, 22050)
    job = engine.CorrelationJob(db, cfg, inp)
    for rep in range(4):
        t = time.time()
        res = job.run()
        wall = time.time() - t
        tm = job.timing()
        n = job.num_offsets
        print(json.dumps(dict(rep=rep, offsets=n, wall_ms=round(wall * 1e3, 3), scan_ms=round(tm["scan_ms"], 3),
                              select_ms=round(tm["select_ms"], 3),
                              scan_offsets_per_s=round(n / (tm["scan_ms"] * 1e-3), 1),
                              scan_tflops=round(n * 4904 / (tm["scan_ms"] * 1e-3) / 1e12, 2),
                              top=res[0] if res else None)), flush=True)


if __name__ == "__main__":
    main()
