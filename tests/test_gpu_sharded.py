"""The sharded-search protocol on ONE GPU: two (or three) database shards driven through
scan -> local_summary -> set_global -> (select -> records -> merge)* exactly like strugatzki_b200/distributed.py
does across ranks, and compared with the single-database result.  (The real 2-GPU run is tools/multi_gpu_check.py.)"""
import numpy as np
import pytest

from util import N, O, STEP, assert_matches_equal, corr_cfgs, make_input, synth

pytestmark = pytest.mark.gpu


def drive(jobs):
    for j in jobs:
        j.scan()
    sums = [j.local_summary() for j in jobs]
    everything = np.concatenate(sums)
    first = np.cumsum([0] + [s.shape[0] for s in sums])
    for j, f0 in zip(jobs, first):
        j.set_global(everything, int(f0))
    done, rounds = False, 0
    while not done:
        recs = np.concatenate([j.select() for j in jobs])
        flags = [j.merge(recs) for j in jobs]
        assert len(set(flags)) == 1
        done = flags[0]
        rounds += 1
        assert rounds < 1000
    res = [j.result() for j in jobs]
    assert all(r == res[0] for r in res), "replicated merge must give identical results on every shard"
    return res[0], rounds


@pytest.mark.parametrize("punch_out,num_matches,num_per_file,split", [
    (False, 100, 1, (70, 140)), (False, 9, 3, (5, 100)), (True, 20, 2, (60, 150)), (True, 40, 1, (100,)),
])
def test_sharded_equals_single(ctx, punch_out, num_matches, num_per_file, split):
    from strugatzki_b200 import engine
    mu, sigma, floor0, norm = synth.default_profile(14)
    # files of 94 x 64 frames: every shard begins on a multiple of 64 frames of the single database, where the sims are
    # bit-identical (the order in which the tensor cores add an offset's products depends on its position modulo 64)
    n_files, frames = 200, 6016
    inp = make_input(900)
    rng = np.random.default_rng(11)
    plants = [(int(rng.integers(0, n_files)), int(rng.integers(0, frames - 800))) for _ in range(15)]
    po = (345 * STEP, 517 * STEP) if punch_out else None
    _, cfg = corr_cfgs(inp, norm, punch_out=po, num_matches=num_matches, num_per_file=num_per_file, min_spacing=22050)

    def build(lo, hi):
        db = engine.Database(ctx, 14, norm)
        for g in range(lo, hi):
            db.add_synth(synth.BASE_SEED, 1 + g, frames, mu, sigma, float(floor0))
        for k, (f, off) in enumerate(plants):
            if lo <= f < hi:
                db.patch(f - lo, off, synth.plant(inp[:172], 3, 2 * k))
                db.patch(f - lo, off + 300, synth.plant(inp[345:517], 3, 2 * k + 1))
        db.finalize()
        return db

    single = engine.CorrelationJob(build(0, n_files), cfg, inp).run()
    bounds = (0,) + tuple(split) + (n_files,)
    shards = [build(a, b) for a, b in zip(bounds, bounds[1:])]
    jobs = [engine.CorrelationJob(s, cfg, inp) for s in shards]
    sharded, rounds = drive(jobs)
    assert sharded == single
    assert rounds >= 2
    found = {(m["file"], m["start"]) for m in sharded}
    if not punch_out and num_matches >= 100:
        assert all((f, off * STEP) in found for f, off in plants)   # every planted needle is recovered


def drive_one_exchange(jobs):
    """what distributed.sharded_search does with job.one_exchange: ONE gather of every shard's best entries"""
    for j in jobs:
        j.scan()
    msgs = [j.local_best() for j in jobs]
    first = np.cumsum([0] + [nf for _, nf, _ in msgs])
    parts = []
    for (recs, _, _), f0 in zip(msgs, first):
        e = recs.copy()
        e["file"] += f0
        parts.append(e)
    best = np.concatenate(parts)
    all_ok = all(ok for _, _, ok in msgs)
    if all_ok:
        for j in jobs:
            j.finish_from_best(best, int(first[-1]))
        rounds = 0
    else:
        entries = np.zeros(best.shape[0], N.ENTRY_DTYPE)
        entries["file"], entries["maxSim"] = best["file"], best["sim"]
        for j, f0 in zip(jobs, first):
            j.set_global_top(entries, int(first[-1]), int(f0))
        done, rounds = False, 0
        while not done:
            recs = np.concatenate([j.select() for j in jobs])
            done = all([j.merge(recs) for j in jobs])
            rounds += 1
            assert rounds < 1000
    res = [j.result() for j in jobs]
    assert all(r == res[0] for r in res)
    return res[0], rounds, all_ok


@pytest.mark.parametrize("num_matches,split,silence", [
    (100, (70, 140), False), (7, (5, 100), False), (250, (100,), False), (30, (60, 150), True), (3, (1, 2, 3, 199), False),
])
def test_one_exchange_equals_single_and_oracle(ctx, num_matches, split, silence):
    """numPerFile = 1, punch-in only: the entries of each shard's numMatches best files decide the search (one exchange,
    no selection kernels).  More matches than files with offsets, shards of a single file, and a shard with digital silence
    (NaN windows: that shard reports ok = False and everybody falls back to the round protocol) -- always the single-GPU
    result, which on a database small enough is the oracle's."""
    from strugatzki_b200 import engine
    mu, sigma, floor0, norm = synth.default_profile(14)
    n_files, frames = 200, 1500
    inp = make_input(900)
    rng = np.random.default_rng(17)
    files = [synth.synth_file(synth.BASE_SEED, 1 + g, frames if g % 17 else 120, mu, sigma, floor0) for g in range(n_files)]
    for k in range(12):
        f, off = int(rng.integers(0, n_files)), int(rng.integers(0, frames - 400))
        if files[f].shape[0] >= frames:
            files[f][off:off + 172] = synth.plant(inp[:172], 3, k)
    if silence:
        files[65][200:700] = files[65][200]          # constant stretch: NaN windows in the second shard
    op, cfg = corr_cfgs(inp, norm, num_matches=num_matches, num_per_file=1, min_spacing=22050)

    def build(lo, hi):
        db = engine.Database(ctx, 14, norm)
        for g in range(lo, hi):
            db.add_file(files[g])
        db.finalize()
        return db

    single = engine.CorrelationJob(build(0, n_files), cfg, inp).run()
    assert_matches_equal(single, O.corr_search(op, files))
    bounds = (0,) + tuple(split) + (n_files,)
    jobs = [engine.CorrelationJob(build(a, b), cfg, inp) for a, b in zip(bounds, bounds[1:])]
    assert all(j.one_exchange for j in jobs)
    sharded, rounds, all_ok = drive_one_exchange(jobs)
    # same files, spans and boosts; the sims agree to the kernel's rounding -- the order in which the tensor cores add an
    # offset's products, and the frame a thread's run of 16 offsets is centred on, depend on the offset's position in the
    # DATABASE modulo 64 (bit-identical when the shards begin on multiples of 64 frames, as in test_sharded_equals_single)
    assert_matches_equal(sharded, single, rel=1e-6)
    assert all_ok == (not silence) and (rounds == 0) == all_ok


def test_search_is_idempotent_and_order_sensitive(ctx):
    """same job run twice -> identical result; the DB order is part of the semantics (Q3): reversing it may change
    which equal-rank matches survive, but never the best match"""
    from strugatzki_b200 import engine
    mu, sigma, floor0, norm = synth.default_profile(14)
    inp = make_input(900)
    files = [synth.synth_file(synth.BASE_SEED, 1 + i, 3000, mu, sigma, floor0) for i in range(20)]
    files[7][500:672] = synth.plant(inp[:172], 9, 1)
    _, cfg = corr_cfgs(inp, norm, num_matches=5, num_per_file=2)

    def run(fs):
        db = engine.Database(ctx, 14, norm)
        for f in fs:
            db.add_file(f)
        db.finalize()
        job = engine.CorrelationJob(db, cfg, inp)
        a, b = job.run(), job.run()
        assert a == b
        return a

    fwd, rev = run(files), run(files[::-1])
    # the needle's score does not depend on where its file sits in the database, up to the kernel's rounding (the
    # tensor-core kernel accumulates an offset's taps in a K order that depends on the offset's position in its tile)
    assert fwd[0]["file"] == 7 and rev[0]["file"] == 12 and abs(fwd[0]["sim"] - rev[0]["sim"]) < 1e-6
    assert fwd[0]["start"] == rev[0]["start"] == 500 * STEP


@pytest.mark.parametrize("frames,decim,world", [(2000, 1, 3), (3000, 4, 2), (2400, 3, 4)])
def test_selfsim_column_blocks_sum_to_full_image(ctx, frames, decim, world):
    """SelfSimilarity shards by blocks of image columns (strugatzki_b200/distributed.py): the partial images of the
    blocks hold disjoint pixels and add up to the image rendered in one go."""
    from strugatzki_b200 import engine
    from strugatzki_b200.distributed import selfsim_column_blocks
    f1, _ = synth.regime_file(synth.BASE_SEED, 21, frames, 14, 6)
    _, _, _, norm = synth.default_profile(14)
    cfg = N.SelfConfig(STEP, 0, 0, 0, 0, 44100, decim, 0.5, 0, 1.0, 1.0, None, 0, 0)
    full, g = engine.self_run(ctx, cfg, f1, None, norm)
    assert engine.self_last_kernel(ctx) == "tc_gram"
    ext = g["imgExt"]
    blocks = selfsim_column_blocks(ext, world)
    total = np.zeros_like(full)
    for b, e in blocks:
        if e > b:
            part, _ = engine.self_run(ctx, cfg, f1, None, norm, b, e)
            assert not np.any((total != 0) & (part != 0) & (total != part))
            total += part
    assert np.array_equal(total, full)
