"""torchrun --nproc-per-node N tools/multi_gpu_check.py : a DB sharded over N GPUs must give EXACTLY the result
of the same DB on one GPU (file-range sharding + all_gather of summaries / candidate records); a SelfSimilarity image
rendered as column blocks on N GPUs must equal the image of one GPU."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from strugatzki_b200 import _native as N, engine, synth  # noqa: E402
from strugatzki_b200.distributed import sharded_search, sharded_self_similarity  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mu, sigma, floor0, norm = synth.default_profile(14)
    n_files, frames = 48 * world, 6000
    inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
    rng = np.random.default_rng(5)
    plants = [(int(rng.integers(0, n_files)), int(rng.integers(0, frames - 700))) for _ in range(12)]
    ctx = engine.Context(local)

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

    per = n_files // world
    shard = build(rank * per, (rank + 1) * per)
    ok = True
    for name, cfg in (
        ("in-only K=100", N.CorrConfig(512, 0, 88200, 0.5, 0, 0, 0, 0.5, 44100, 352800, 8.0, 100, 1, 22050)),
        ("in-only K=7 npf=3", N.CorrConfig(512, 0, 88200, 0.5, 0, 0, 0, 0.5, 44100, 352800, 8.0, 7, 3, 0)),
        ("in+out K=20 npf=2", N.CorrConfig(512, 0, 88200, 0.5, 1, 176640, 264704, 0.5, 44100, 352800, 8.0, 20, 2, 22050)),
    ):
        res = sharded_search(engine.CorrelationJob(shard, cfg, inp), dev, None)
        if rank == 0:
            full = build(0, n_files)
            ref = engine.CorrelationJob(full, cfg, inp).run()
            same = res == ref
            ok &= same
            print(f"{name}: {len(res)} matches, sharded x{world} == single GPU: {same}; top {res[0]}", flush=True)
            full.close()
    # SelfSimilarity: column blocks of the image, one per GPU, summed on rank 0 == the image rendered by one GPU
    seg, _ = synth.regime_file(synth.BASE_SEED, 21, 6000, 14, 8)
    scfg = N.SelfConfig(512, 0, 0, 0, 0, 44100, 1, 0.5, 0, 1.0, 1.0, None, 0, 0)
    ext = engine.self_geometry(scfg, seg.shape[0], seg.shape[0])["imgExt"]
    img = sharded_self_similarity(lambda b, e: engine.self_run(ctx, scfg, seg, None, norm, b, e)[0], ext, dev, None)
    if rank == 0:
        ref, _ = engine.self_run(ctx, scfg, seg, None, norm)
        same = bool(np.array_equal(img, ref))
        ok &= same
        print(f"self-similarity {ext} x {ext} ({engine.self_last_kernel(ctx)}): column blocks x{world} == single GPU: {same}", flush=True)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
