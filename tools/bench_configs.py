"""BASELINE.json configs[0..3] on one GPU with the oracle timed beside them (developer tool, writes JSON lines).

  configs[0]  FeatureCorrelation: 2 s punch-in vs 100 x 10 min files, numMatches 10        (full parity vs oracle)
  configs[1]  FeatureSegmentation: 10 min file, corrLen 0.5 s, 20 breaks                   (bit-identical breaks)
  configs[2]  FeatureCorrelation punch-in + punch-out, 1 s / 8 s, 10 h DB (60 files)       (parity vs oracle)
  configs[3]  SelfSimilarity of a ~155 000-frame file -> 38 707^2 image                     (cells/s; parity on a sample)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402  (developer tool: oracle used as checker / CPU baseline)
from strugatzki_b200 import _native as N, engine, synth  # noqa: E402

STEP, FR = 512, 51680
quick = "--quick" in sys.argv
mu, sigma, floor0, norm = synth.default_profile(14)
inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
ctx = engine.Context(0)


def same(got, want):
    return len(got) == len(want) and all(
        (g["file"], g["start"], g["stop"]) == (w["file"], w["start"], w["stop"]) and
        abs(g["sim"] - w["sim"]) <= 1e-5 * abs(w["sim"]) + 2e-6 for g, w in zip(got, want))


def corr_case(name, n_files, cfg_kw, plants, cpu_files):
    files = [synth.synth_file(synth.BASE_SEED, 1 + i, FR, mu, sigma, floor0) for i in range(n_files)]
    for k, (f, a, b) in enumerate(plants):
        files[f][a:a + 172] = synth.plant(inp[:172], 41, 2 * k)
        if b is not None:
            files[f][b:b + 172] = synth.plant(inp[345:517], 41, 2 * k + 1)
    db = engine.Database(ctx, 14, norm)
    for f in files:
        db.add_file(f)
    db.finalize()
    po = cfg_kw.get("punch_out")
    cfg = N.CorrConfig(STEP, 0, 88200, 0.5, 0 if po is None else 1, 0 if po is None else po[0], 0 if po is None else po[1],
                       0.5, cfg_kw["min_punch"], cfg_kw["max_punch"], 8.0, cfg_kw["num_matches"], cfg_kw["num_per_file"],
                       cfg_kw["min_spacing"])
    job = engine.CorrelationJob(db, cfg, inp)
    job.run()
    times = []
    for _ in range(5):
        ctx.synchronize(); t = time.perf_counter(); got = job.run(); ctx.synchronize(); times.append(time.perf_counter() - t)
    gpu_s = float(np.median(times))
    op = O.CorrParams(step_size=STEP, input=inp, punch_in=(0, 88200), punch_out=po, min_punch=cfg_kw["min_punch"],
                      max_punch=cfg_kw["max_punch"], norm=norm, num_matches=cfg_kw["num_matches"],
                      num_per_file=cfg_kw["num_per_file"], min_spacing=cfg_kw["min_spacing"])
    sub = files[:cpu_files]
    t = time.perf_counter(); want_sub = O.corr_search(op, sub); cpu_s = time.perf_counter() - t
    n_sub = O.corr_num_offsets(op, [f.shape[0] for f in sub])
    parity = None
    if cpu_files == n_files:
        parity = same(got, want_sub)
    else:   # oracle on the prefix must equal the engine on the same prefix
        db2 = engine.Database(ctx, 14, norm)
        for f in sub:
            db2.add_file(f)
        db2.finalize()
        parity = same(engine.CorrelationJob(db2, cfg, inp).run(), want_sub)
    tm = job.timing()
    print(json.dumps(dict(config=name, files=n_files, offsets=job.num_offsets, gpu_ms=round(gpu_s * 1e3, 3),
                          scan_ms=round(tm["scan_ms"], 3), select_ms=round(tm["select_ms"], 3),
                          gpu_offsets_per_s=round(job.num_offsets / gpu_s, 1), matches=len(got), top=got[0],
                          cpu_oracle_files=cpu_files, cpu_oracle_s=round(cpu_s, 2),
                          cpu_offsets_per_s=round(n_sub / cpu_s, 1), parity_vs_oracle=parity)), flush=True)


corr_case("configs[0] FeatureCorrelation 2 s punch-in, 100 x 10 min files, numMatches 10", 100 if not quick else 10,
          dict(min_punch=44100, max_punch=352800, num_matches=10, num_per_file=1, min_spacing=0),
          [(3, 1000, None), (57 if not quick else 7, 40000, None)], 100 if not quick else 10)

# configs[1] segmentation
seg, cuts = synth.regime_file(synth.BASE_SEED, 31, FR, 14, 26)
scfg = N.SegmConfig(STEP, 0, 0, 0, 0, 22050, 0.5, 20, 22050)
engine.segm_run(ctx, scfg, seg, norm)
t = time.perf_counter(); gb, curve, noff = engine.segm_run(ctx, scfg, seg, norm, want_curve=True); g_s = time.perf_counter() - t
kms, _ = ctx.last_timing()
t = time.perf_counter(); wb, wcurve = O.segm_run(O.SegmParams(step_size=STEP, norm=norm, num_breaks=20), seg, want_curve=True)
c_s = time.perf_counter() - t
ident = [(b["pos"], np.float32(b["sim"]).tobytes()) for b in gb] == [(b["pos"], np.float32(b["sim"]).tobytes()) for b in wb]
print(json.dumps(dict(config="configs[1] FeatureSegmentation 10 min, corrLen 0.5 s, 20 breaks", offsets=int(noff),
                      gpu_wall_ms=round(g_s * 1e3, 3), gpu_kernels_ms=round(kms, 3), cpu_oracle_s=round(c_s, 3),
                      breaks_bit_identical=ident, curve_bit_identical=bool(np.array_equal(
                          curve.view(np.uint32), wcurve[:noff].view(np.uint32))),
                      breaks_near_planted_cuts=int(sum(min(abs(b["pos"] // STEP - c) for c in cuts) <= 3 for b in gb)))),
      flush=True)

corr_case("configs[2] FeatureCorrelation punch-in+out, minPunch 1 s / maxPunch 8 s, 10 h DB", 60 if not quick else 8,
          dict(punch_out=(176640, 264704), min_punch=44100, max_punch=352800, num_matches=20, num_per_file=2,
               min_spacing=22050), [(2, 5000, 5400), (33 if not quick else 5, 30000, 30650)], 12 if not quick else 4)

# configs[3] self similarity
n = 155000 if not quick else 12000
sf, _ = synth.regime_file(synth.BASE_SEED, 32, n, 14, 60)
cfg = N.SelfConfig(STEP, 0, 0, 0, 0, 44100, 1, 0.5, 0, 1.0, 1.0, None, 0)
t = time.perf_counter(); _, g = engine.self_run(ctx, cfg, sf, None, norm, download=False); wall = time.perf_counter() - t
kms, _ = ctx.last_timing()
rng = np.random.default_rng(2)
l = rng.integers(0, g["imgExt"], 2000); r = rng.integers(0, g["imgExt"], 2000)
gs, grgb = engine.self_cells(ctx, cfg, sf, None, l, r, norm)
t = time.perf_counter(); ws, wrgb = O.self_cells(O.SelfParams(step_size=STEP, corr_len=44100, decimation=1, norm=norm), sf, None, l, r)
c_s = time.perf_counter() - t
print(json.dumps(dict(config="configs[3] SelfSimilarity ~155k frames", frames=n, imgExt=g["imgExt"], decim=g["decim"],
                      cells=g["numCells"], gpu_kernel_ms=round(kms, 2), gpu_wall_ms=round(wall * 1e3, 1),
                      gpu_cells_per_s=round(g["numCells"] / (kms * 1e-3), 1), cpu_oracle_cells_per_s=round(2000 / c_s, 1),
                      mode="tensor-core Gram (default)",
                      sample_max_abs_err=float(np.max(np.abs(gs - ws))),
                      sample_max_rel_err_where_abs_sim_gt_0p05=float(np.max((np.abs(gs - ws) / np.abs(ws))[np.abs(ws) > 0.05])),
                      sample_max_grey_diff=int(np.max(np.abs((grgb & 255) - (wrgb & 255)))))), flush=True)
pcfg = N.SelfConfig(STEP, 0, 0, 0, 0, 44100, 1, 0.5, 0, 1.0, 1.0, None, 0, 1)
t = time.perf_counter(); _, g = engine.self_run(ctx, pcfg, sf, None, norm, download=False); wall = time.perf_counter() - t
kms, _ = ctx.last_timing()
gs, grgb = engine.self_cells(ctx, pcfg, sf, None, l, r, norm)
print(json.dumps(dict(config="configs[3] SelfSimilarity ~155k frames", mode="precise FP64 replay", cells=g["numCells"],
                      gpu_kernel_ms=round(kms, 2), gpu_cells_per_s=round(g["numCells"] / (kms * 1e-3), 1),
                      sample_cells_bit_identical=bool(np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
                                                      and np.array_equal(grgb, wrgb)))), flush=True)


# ---- section 8(f) rows: CrossSimilarity and FeatureStats ----
cf = [synth.synth_file(synth.BASE_SEED, 900, 172, mu, sigma, floor0),
      synth.synth_file(synth.BASE_SEED, 901, 310078 if not quick else 20000, mu, sigma, floor0)]
ccfg = N.CrossConfig(STEP, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.5, 8.0)
engine.cross_run(ctx, ccfg, cf[0], cf[1], norm)
t = time.perf_counter(); gc = engine.cross_run(ctx, ccfg, cf[0], cf[1], norm); wall = time.perf_counter() - t
kms, _ = ctx.last_timing()
sub = cf[1][:8192 + 20000]
t = time.perf_counter(); wc = O.cross_run(O.CrossParams(step_size=STEP, norm=norm), cf[0], sub); c_s = time.perf_counter() - t
print(json.dumps(dict(config="8(f) CrossSimilarity: 2 s template over a 1 h feature file", outputs=int(gc.shape[0]),
                      gpu_kernel_ms=round(kms, 3), gpu_wall_ms=round(wall * 1e3, 1),
                      gpu_outputs_per_s=round(gc.shape[0] / (kms * 1e-3), 1),
                      cpu_oracle_outputs_per_s=round(wc.shape[0] / c_s, 1),
                      prefix_bit_identical=bool(np.array_equal(gc[:wc.shape[0]].view(np.uint32), wc.view(np.uint32))))),
      flush=True)

nst = 600 if not quick else 20
raw = engine.Database(ctx, 14, None)
raw.reserve(nst * 51680, nst)
for i in range(nst):
    raw.add_synth(synth.BASE_SEED, 5000 + i, 51680, mu, sigma, float(floor0))
raw.finalize()
raw.stats()
t = time.perf_counter(); gst, gper = raw.stats(want_per_file=True); wall = time.perf_counter() - t
kms, _ = ctx.last_timing()
ncpu = 6
fl = [synth.synth_file(synth.BASE_SEED, 5000 + i, 51680, mu, sigma, floor0) for i in range(ncpu)]
t = time.perf_counter(); _, wper = O.stats_run(fl, want_per_file=True); c_s = time.perf_counter() - t
frames = nst * 51680
print(json.dumps(dict(config="8(f) FeatureStats over a 100 h raw feature database", files=nst, frames=frames,
                      gpu_kernels_ms=round(kms, 3), gpu_wall_ms=round(wall * 1e3, 1),
                      gpu_frames_per_s=round(frames / (kms * 1e-3), 1),
                      algorithmic_GBps=round(frames * 112 / (kms * 1e-3) / 1e9, 1),
                      cpu_oracle_frames_per_s=round(ncpu * 51680 / c_s, 1),
                      per_file_max_rel_err=float(np.max(np.abs(gper[:ncpu] - wper) / np.abs(wper))))), flush=True)
