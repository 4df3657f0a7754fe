"""N > 1 host logic on CPU: gloo, world_size 2.  The sharded-search driver only moves bytes between ranks and
calls the job's phase methods in lock-step; here the job is a pure-Python stand-in that follows the same
protocol, so the control flow (rank-order concatenation, first-file offsets, termination) is covered without
a GPU."""
import os
import sys

import numpy as np
import pytest

from util import N, ROOT


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from strugatzki_b200.distributed import allgather_bytes, sharded_search
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. ragged all-gather of POD records, rank order preserved, empty contributions allowed
        mine = np.zeros(3 * rank, N.RECORD_DTYPE)
        mine["file"] = rank
        mine["piOff"] = np.arange(3 * rank)
        for _ in range(2):   # 1st call overflows the tiny capacity (two collectives), 2nd fits in one
            got, cnt = allgather_bytes(mine, tag="t1", initial_cap=32)
            assert got.dtype == N.RECORD_DTYPE and got.shape[0] == sum(3 * r for r in range(world))
            assert cnt == [3 * r for r in range(world)]
            assert list(got["file"]) == [r for r in range(world) for _ in range(3 * r)]
            assert list(got["piOff"]) == [i for r in range(world) for i in range(3 * r)]

        # 2. protocol driver with a stand-in job: each rank owns (rank + 2) files whose maximum is known
        class FakeJob:
            def __init__(self):
                self.n_local = rank + 2
                self.calls = []
                self.all = None
                self.rounds = 0
                self.best = []

            def scan(self):
                self.calls.append("scan")

            def local_summary(self):
                s = np.zeros(self.n_local, N.SUMMARY_DTYPE)
                s["maxSim"] = [0.1 * (rank + 1) + 0.01 * i for i in range(self.n_local)]
                s["numOffsets"] = 100
                return s

            def set_global(self, everything, my_first):
                self.all, self.first = everything.copy(), my_first

            def select(self):
                # round 0: every rank reports its files' maxima; round 1: nothing left
                if self.rounds == 0:
                    r = np.zeros(self.n_local, N.RECORD_DTYPE)
                    r["file"] = self.first + np.arange(self.n_local)
                    r["sim"] = self.all["maxSim"][self.first:self.first + self.n_local]
                    return r
                return np.zeros(0, N.RECORD_DTYPE)

            def merge(self, recs):
                self.rounds += 1
                if recs.shape[0]:
                    assert list(recs["file"]) == sorted(recs["file"])       # rank order == file order
                    self.best = sorted(((float(s), int(f)) for s, f in zip(recs["sim"], recs["file"])),
                                       reverse=True)[:3]
                return self.rounds == 2

            def result(self):
                return [dict(sim=s, file=f) for s, f in self.best]

        job = FakeJob()
        res = sharded_search(job)
        total = sum(r + 2 for r in range(world))
        assert job.all.shape[0] == total and job.first == sum(r + 2 for r in range(rank))
        assert job.rounds == 2 and job.calls == ["scan"]

        # 3. the sparse summary of punch-in-only searches: every rank sends its file count and its top entries with LOCAL
        # indices; the driver rebases them to the global file list
        class SparseJob(FakeJob):
            sparse_summary = True

            def local_top(self):
                e = np.zeros(1, N.ENTRY_DTYPE)              # only the best local file
                e["file"] = self.n_local - 1
                e["maxSim"] = 0.1 * (rank + 1) + 0.01 * (self.n_local - 1)
                return e, self.n_local

            def set_global_top(self, entries, n_files_global, my_first):
                self.entries, self.n_global, self.first = entries.copy(), n_files_global, my_first
                self.all = np.zeros(n_files_global, N.SUMMARY_DTYPE)
                self.all["maxSim"] = -np.inf
                self.all["maxSim"][entries["file"]] = entries["maxSim"]

            def local_summary(self):
                raise AssertionError("the dense summary must not be used")

        sj = SparseJob()
        res2 = sharded_search(sj)
        assert sj.n_global == total and sj.first == sum(r + 2 for r in range(rank))
        assert list(sj.entries["file"]) == [sum(q + 2 for q in range(r)) + r + 1 for r in range(world)]
        assert np.allclose(sj.entries["maxSim"], [0.1 * (r + 1) + 0.01 * (r + 1) for r in range(world)])
        # 4. one match per file: ONE exchange of every rank's best entries; rank 1 once reports a NaN file (ok = False) and
        # everybody must then continue with the round protocol on the same gathered bytes
        class BestJob(SparseJob):
            one_exchange = True
            nan_rank = -1

            def local_best(self):
                r = np.zeros(1, N.RECORD_DTYPE)
                r["file"], r["kind"], r["piOff"] = self.n_local - 1, 1, 7 + rank
                r["sim"] = 0.1 * (rank + 1) + 0.01 * (self.n_local - 1)
                return r, self.n_local, rank != self.nan_rank

            def finish_from_best(self, best, n_files_global):
                self.finished_with = best.copy()
                self.n_global = n_files_global
                self.best = sorted(((float(x), int(f)) for x, f in zip(best["sim"], best["file"])), reverse=True)[:3]

            def local_top(self):
                raise AssertionError("the top entries travel inside the best records")

        bj = BestJob()
        res3 = sharded_search(bj)
        assert bj.rounds == 0 and bj.n_global == total                       # no selection round at all
        assert list(bj.finished_with["file"]) == [sum(q + 2 for q in range(r)) + r + 1 for r in range(world)]
        assert list(bj.finished_with["piOff"]) == [7 + r for r in range(world)]
        fb = BestJob()
        fb.nan_rank = 1
        sharded_search(fb)
        assert not hasattr(fb, "finished_with") and fb.rounds == 2            # fell back: set_global_top + rounds
        assert list(fb.entries["file"]) == list(bj.finished_with["file"])
        assert np.allclose(fb.entries["maxSim"], bj.finished_with["sim"])
        out_q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_protocol():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results[0] == results[1]                                          # replicated merge -> identical result
    assert [m["file"] for m in results[0]] == [4, 3, 2]


def _selfsim_worker(rank, world, port, q):
    import os
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from strugatzki_b200.distributed import selfsim_column_blocks, sharded_self_similarity
    ext = 700
    full = np.zeros((ext, ext), np.int32)
    a, c = np.triu_indices(ext)                       # cell (a, c >= a): pixel (ext-1-c, a) and mirror (ext-1-a, c)
    val = ((a * 131 + c * 7) % 255 + 1).astype(np.int32) * 0x010101
    full[ext - 1 - c, a] = val
    full[ext - 1 - a, c] = val

    def render(b, e):                                 # what sgz_self_run(rowBegin = b, rowEnd = e) leaves in the image
        img = np.zeros((ext, ext), np.int32)
        m = (a >= b) & (a < e)
        img[ext - 1 - c[m], a[m]] = val[m]
        img[ext - 1 - a[m], c[m]] = val[m]
        return img

    got = sharded_self_similarity(render, ext)
    blocks = selfsim_column_blocks(ext, world)
    q.put((rank, None if got is None else bool(np.array_equal(got, full)), blocks))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_self_similarity_world2():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 17
    ps = [ctx.Process(target=_selfsim_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] is True and res[1][1] is None
    assert res[0][2] == res[1][2]


def test_selfsim_column_blocks_balanced():
    from strugatzki_b200.distributed import selfsim_column_blocks
    for ext, world in ((38707, 8), (29829, 4), (700, 2), (100, 8), (1, 2), (129, 3)):
        blocks = selfsim_column_blocks(ext, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == ext
        assert all(b0[1] == b1[0] for b0, b1 in zip(blocks, blocks[1:]))
        assert all(b % 128 == 0 or b == ext for blk in blocks for b in blk)
        if ext > 128 * 8 * world:
            cells = [sum(ext - a for a in range(b, e)) for b, e in blocks]
            assert max(cells) < 1.15 * (sum(cells) / world), (ext, world, cells)
