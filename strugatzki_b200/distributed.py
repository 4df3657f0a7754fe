"""Sharded FeatureCorrelation search: one process per GPU, torch.distributed for the plumbing.

The database shards by contiguous ranges of the (ordered) file list -- rank r holds the files after all
files of ranks < r.  The arithmetic (K1) needs no communication at all.  The only exchange is the
reference's sequential coupling between files, `(allPrio.size, allPrio.last.sim)`
(FeatureCorrelationImpl.scala:120-129, 399-400): per-file maxima once, then a few KB of candidate records
per selection round, moved with all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests).  Every rank
then runs the same deterministic replay (sgz_corr_merge), so all ranks end with the identical result and no
broadcast is needed.

The payloads are KBs, so the cost is collective LATENCY: every exchange is ONE fixed-capacity all_gather
whose first 8 bytes carry the true length (a second, larger collective only when some rank overflowed; the
capacity then grows for the following searches -- identically on every rank).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import numpy as np

_caps: Dict[str, int] = {}
_bufs: Dict[tuple, tuple] = {}


def _exchange_buffers(tag: str, cap: int, world: int, dev):
    """Persistent buffers of one exchange: pinned host send / receive staging and the device send / receive tensors NCCL
    works on, allocated once per (tag, capacity) -- a search runs the same two exchanges every time, and a fresh pageable
    tensor per call costs a synchronous staged copy in each direction."""
    import torch
    key = (tag, cap, world, str(dev))
    b = _bufs.get(key)
    if b is None:
        n = 8 + cap
        if dev.type == "cuda":
            b = (torch.zeros(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, device=dev),
                 torch.empty(world * n, dtype=torch.uint8, device=dev), torch.empty(world * n, dtype=torch.uint8, pin_memory=True))
        else:
            b = (torch.zeros(n, dtype=torch.uint8), None, None, torch.empty(world * n, dtype=torch.uint8))
        for k in [k for k in _bufs if k[0] == tag and k[2:] == key[2:]]:     # an outgrown capacity is not kept
            del _bufs[k]
        _bufs[key] = b
    return b


def allgather_bytes(arr: np.ndarray, device=None, group=None, tag: str = "default", initial_cap: int = 1 << 16):
    """All-gather a 1-D structured/POD numpy array of rank-dependent length.  Returns (concatenation in rank
    order, list of per-rank element counts).  Works with NCCL (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dtype = arr.dtype
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    dev = device if device is not None else torch.device("cpu")
    cap = _caps.get(tag, initial_cap)
    while True:
        h_send, d_send, d_recv, h_recv = _exchange_buffers(tag, cap, world, dev)
        msg = h_send.numpy()
        msg[:8] = np.frombuffer(np.int64(raw.shape[0]).tobytes(), np.uint8)
        n_send = min(raw.shape[0], cap)
        msg[8:8 + n_send] = raw[:n_send]
        if dev.type == "cuda":
            d_send.copy_(h_send, non_blocking=True)
            dist.all_gather_into_tensor(d_recv, d_send, group=group)
            h_recv.copy_(d_recv, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        else:
            dist.all_gather(list(h_recv.view(world, 8 + cap).unbind(0)), h_send, group=group)
        host = h_recv.numpy().reshape(world, 8 + cap)
        sizes = [int(np.frombuffer(host[r, :8].tobytes(), np.int64)[0]) for r in range(world)]
        if max(sizes) <= cap:
            break
        cap = 2 * max(sizes)          # same decision on every rank
        _caps[tag] = cap
    parts = [host[r, 8:8 + sizes[r]] for r in range(world)]
    cat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return cat.view(dtype), [s // dtype.itemsize for s in sizes]


def sharded_search(job, device=None, group=None, trace: Optional[dict] = None) -> List[dict]:
    """Run one search over a database sharded across the ranks of `group`; `job` is this rank's
    CorrelationJob (or any object with the same scan/summary/select/merge surface).  `trace`, when given,
    accumulates host wall seconds per phase (every phase ends with a device synchronisation)."""
    import torch.distributed as dist

    t_last = time.perf_counter()

    def mark(phase):
        nonlocal t_last
        if trace is not None:
            now = time.perf_counter()
            trace[phase] = trace.get(phase, 0.0) + (now - t_last)
            t_last = now

    rank = dist.get_rank(group)
    job.scan()
    mark("scan")
    if getattr(job, "one_exchange", False):
        # punch-in only, one match per file: the entries of every rank's numMatches best files are all the search needs
        # (strugatzki_b200.h, sgz_corr_local_best); record 0 of a rank's message carries its number of files and whether
        # the shortcut applies there.  Every rank sees the same gathered bytes, so all take the same branch.
        from . import _native as N
        recs, n_local, ok = job.local_best()
        msg = np.zeros(recs.shape[0] + 1, N.RECORD_DTYPE)
        msg[0]["file"] = n_local
        msg[0]["kind"] = 1 if ok else 0
        msg[1:] = recs
        mark("summary_download")
        everything, counts = allgather_bytes(msg, device, group, tag="best", initial_cap=1 << 13)
        mark("summary_allgather")
        cnt = np.asarray(counts, np.int64)
        starts = np.concatenate([[0], np.cumsum(cnt)])
        hdr = everything[starts[:-1]]                           # record 0 of every rank
        all_ok = bool((hdr["kind"] == 1).all())
        first = np.concatenate([[0], np.cumsum(hdr["file"].astype(np.int64))])
        keep = np.ones(everything.shape[0], bool)
        keep[starts[:-1]] = False
        best = everything[keep]                                 # (a copy) all entries, rank order = file order
        best["file"] += first[np.repeat(np.arange(cnt.shape[0]), cnt - 1)].astype(np.int32)
        if all_ok:
            job.finish_from_best(best, int(first[-1]))
            mark("merge")
            if trace is not None:
                trace["rounds"] = trace.get("rounds", 0)
            res = job.result()
            mark("result")
            return res
        entries = np.zeros(best.shape[0], N.ENTRY_DTYPE)       # some rank holds NaN windows: the round protocol
        entries["file"] = best["file"]
        entries["maxSim"] = best["sim"]
        job.set_global_top(entries, int(first[-1]), int(first[rank]))
    elif getattr(job, "sparse_summary", False):
        # punch-in only: the numMatches largest file maxima of every rank are all the thresholds need (strugatzki_b200.h);
        # entry 0 of each rank's message carries its number of files
        from . import _native as N
        top, n_local = job.local_top()
        msg = np.zeros(top.shape[0] + 1, N.ENTRY_DTYPE)
        msg[0]["file"] = n_local
        msg[1:] = top
        mark("summary_download")
        everything, counts = allgather_bytes(msg, device, group, tag="top", initial_cap=1 << 12)
        mark("summary_allgather")
        starts = np.concatenate([[0], np.cumsum(counts)])
        n_files = [int(everything[starts[r]]["file"]) for r in range(len(counts))]
        first = np.concatenate([[0], np.cumsum(n_files)])
        parts = []
        for r in range(len(counts)):
            e = everything[starts[r] + 1:starts[r + 1]].copy()
            e["file"] += first[r]
            parts.append(e)
        job.set_global_top(np.concatenate(parts), int(first[-1]), int(first[rank]))
    else:
        summary = job.local_summary()
        mark("summary_download")
        everything, counts = allgather_bytes(summary, device, group, tag="summary", initial_cap=1 << 17)
        mark("summary_allgather")
        job.set_global(everything, int(sum(counts[:rank])))
    mark("set_global")
    done = False
    rounds = 0
    while not done:
        recs = job.select()
        mark("select")
        all_recs, _ = allgather_bytes(recs, device, group, tag="records", initial_cap=1 << 16)
        mark("records_allgather")
        done = job.merge(all_recs)
        mark("merge")
        rounds += 1
        if rounds > 1_000_000:
            raise RuntimeError("selection protocol did not terminate")
    if trace is not None:
        trace["rounds"] = trace.get("rounds", 0) + rounds
    res = job.result()
    mark("result")
    return res


# ---------------------------------------------------------------------------------------------
# SelfSimilarity: the matrix shards by blocks of image columns
# ---------------------------------------------------------------------------------------------
def selfsim_column_blocks(img_ext: int, world: int, tile: int = 128) -> List[tuple]:
    """Column blocks [begin, end) of the upper triangle, one per rank, aligned to the kernel's tile size and balanced by
    CELL count (column a holds img_ext - a cells, so the first blocks are narrower).  Blocks may be empty when the image
    has fewer tiles than ranks."""
    n_tiles = (img_ext + tile - 1) // tile
    cells = np.array([sum(img_ext - a for a in range(t * tile, min((t + 1) * tile, img_ext))) for t in range(n_tiles)], np.float64)
    cum = np.concatenate([[0.0], np.cumsum(cells)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target))            # first boundary at or beyond the target
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n_tiles)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n_tiles))
    cuts.append(n_tiles)
    return [(min(a * tile, img_ext), min(b * tile, img_ext)) for a, b in zip(cuts, cuts[1:])]


def sharded_self_similarity(render, img_ext: int, device=None, group=None, dst: int = 0):
    """One SelfSimilarity image rendered by all ranks of `group`: rank r calls `render(col_begin, col_end)` -- e.g.
    `lambda b, e: engine.self_run(ctx, cfg, frames1, frames2, norm, b, e)[0]` with the feature file replicated on every
    GPU -- for its column block and gets an (img_ext, img_ext) int32 image that holds only that block's pixels and their
    mirrors.  A block [b, e) touches two strips of the image and nothing else: its cells (a, c >= a) sit at
    (row ext-1-c, column a), i.e. in COLUMNS [b, e), and their mirrors at (row ext-1-a, column c), i.e. in ROWS
    [ext-e, ext-b).  Only these two strips travel to rank `dst` (point to point; NCCL over NVLink on GPUs, gloo in the CPU
    tests), 2 x ext x (e - b) pixels per rank -- 8 ext^2 bytes in total whatever the number of ranks, where a reduce of whole
    images moved 4 ext^2 bytes per rank and needed each of them on the device.  Returns the image on `dst`, None elsewhere.
    No other communication: the Gram tiles are independent (SelfSimilarityImpl.scala:127-155 has no cross-cell state)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = selfsim_column_blocks(img_ext, world)
    b, e = blocks[rank]
    dev = device if device is not None else torch.device("cpu")
    part = None
    if e > b:
        part = np.ascontiguousarray(render(b, e), np.int32)
        assert part.shape == (img_ext, img_ext), part.shape
    if rank != dst:
        if part is not None:
            for strip in (part[:, b:e], part[img_ext - e:img_ext - b, :]):
                dist.send(torch.from_numpy(np.ascontiguousarray(strip)).to(dev), dst, group=group)
        return None
    full = part if part is not None else np.zeros((img_ext, img_ext), np.int32)
    for r, (rb, re) in enumerate(blocks):
        if r == dst or re <= rb:
            continue
        v = torch.empty((img_ext, re - rb), dtype=torch.int32, device=dev)
        h = torch.empty((re - rb, img_ext), dtype=torch.int32, device=dev)
        dist.recv(v, r, group=group)
        dist.recv(h, r, group=group)
        full[:, rb:re] |= v.cpu().numpy()           # strips of different ranks cross, but no pixel is written twice
        full[img_ext - re:img_ext - rb, :] |= h.cpu().numpy()
    return full
