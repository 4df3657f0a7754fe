"""Sharded FeatureCorrelation search: one process per GPU, torch.distributed for the plumbing.

The database shards by contiguous ranges of the (ordered) file list -- rank r holds the files after all
files of ranks < r.  The arithmetic (K1) needs no communication at all.  The only exchange is the
reference's sequential coupling between files, `(allPrio.size, allPrio.last.sim)`
(FeatureCorrelationImpl.scala:120-129, 399-400): per-file maxima once, then a few KB of candidate records
per selection round, moved with all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests).  Every rank
then runs the same deterministic replay (sgz_corr_merge), so all ranks end with the identical result and no
broadcast is needed.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


def allgather_bytes(arr: np.ndarray, device=None, group=None) -> np.ndarray:
    """All-gather a 1-D structured/POD numpy array of rank-dependent length; returns the concatenation in
    rank order.  Works with NCCL (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dtype = arr.dtype
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([raw.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if raw.shape[0]:
        buf[:raw.shape[0]] = torch.from_numpy(raw.copy()).to(dev)
    out = torch.zeros(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group) if dev.type == "cuda" else \
        dist.all_gather(list(out.view(world, mx).unbind(0)), buf, group=group)
    host = out.cpu().numpy().reshape(world, mx)
    parts = [host[r, :sizes[r]] for r in range(world)]
    cat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return cat.view(dtype)


def sharded_search(job, device=None, group=None) -> List[dict]:
    """Run one search over a database sharded across the ranks of `group`; `job` is this rank's
    CorrelationJob (or any object with the same scan/summary/select/merge surface)."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    job.scan()
    local = job.local_summary()
    counts = allgather_bytes(np.array([local.shape[0]], np.int64), device, group)
    my_first = int(counts[:rank].sum())
    everything = allgather_bytes(local, device, group)
    job.set_global(everything, my_first)
    done = False
    rounds = 0
    while not done:
        recs = job.select()
        all_recs = allgather_bytes(recs, device, group)
        done = job.merge(all_recs)
        rounds += 1
        if rounds > 1_000_000:
            raise RuntimeError("selection protocol did not terminate")
    return job.result()
