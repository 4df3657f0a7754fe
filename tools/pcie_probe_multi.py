"""Host -> device bandwidth of N GPUs at once (the ceiling of the end-to-end arm at N ranks).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py

Every rank copies a pinned 4 GiB buffer to its GPU: alone (ranks take turns), all at once with the default memory policy,
and all at once with the buffer bound to each NUMA node the container is allowed to use (set_mempolicy, no libnuma
needed).  Prints one JSON line: per-rank and aggregate GB/s, the CPU / memory nodes the process may use, the GPU topology.
"""
import ctypes
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def status_field(name):
    for line in open("/proc/self/status"):
        if line.startswith(name):
            return line.split(":", 1)[1].strip()
    return None


def set_mempolicy(mode, nodes):
    """mode 0 = default, 2 = MPOL_BIND; returns errno (0 = ok)"""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(sum(1 << n for n in nodes))
    rc = libc.syscall(238, ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(64))
    return 0 if rc == 0 else ctypes.get_errno()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


SIZE = 4 << 30
dst = torch.empty(SIZE, dtype=torch.uint8, device=dev)


def measure(buf, together):
    """GB/s of this rank; together = all ranks copy at the same time, else one rank after the other"""
    res = 0.0
    for turn in range(1 if together else world):
        barrier()
        if together or turn == rank:
            dst.copy_(buf, non_blocking=True)       # warm-up
            torch.cuda.synchronize()
        barrier()
        if together or turn == rank:
            t = time.perf_counter()
            for _ in range(3):
                dst.copy_(buf, non_blocking=True)
            torch.cuda.synchronize()
            res = 3 * SIZE / (time.perf_counter() - t) / 1e9
    barrier()
    return res


def gather(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o[0]) for o in out]
    return [x]


nodes_online = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
out = {"world": world, "cpus_allowed": status_field("Cpus_allowed_list"), "mems_allowed": status_field("Mems_allowed_list"),
       "numa_nodes_online": nodes_online}
buf = torch.empty(SIZE, dtype=torch.uint8, pin_memory=True)
buf.zero_()
alone = gather(measure(buf, False))
together = gather(measure(buf, True))
out["default_policy"] = {"alone_GBps": alone, "together_GBps": together, "together_sum_GBps": sum(together)}
del buf
for node in nodes_online:
    err = set_mempolicy(2, [node])
    entry = {"set_mempolicy_errno": err}
    b2 = None
    if err == 0:
        try:
            b2 = torch.empty(SIZE, dtype=torch.uint8, pin_memory=True)
            b2.zero_()
        except Exception as e:      # the container's memory cgroup may not include the node
            entry["error"] = repr(e)[:200]
            b2 = None
    ok = gather(1.0 if b2 is not None else 0.0)      # every rank takes the same branch (barriers inside measure)
    if all(v > 0 for v in ok):
        tg = gather(measure(b2, True))
        entry.update({"together_GBps": tg, "together_sum_GBps": sum(tg)})
    else:
        entry["ranks_that_could_allocate"] = ok
    del b2
    set_mempolicy(0, [])
    out[f"bound_to_node{node}"] = entry
if rank == 0:
    try:
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception as e:
        out["topo"] = repr(e)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
