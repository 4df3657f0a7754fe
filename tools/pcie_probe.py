"""Host->device copy bandwidth from pinned memory: one large copy vs many file-sized copies (sizing the e2e bound)."""
import json
import time

import torch

dev = torch.device("cuda", 0)
out = {}
big = torch.empty(2 << 30, dtype=torch.uint8, pin_memory=True)
big.zero_()
dst = torch.empty_like(big, device=dev)
for name, chunk in (("one_2GiB_copy", 2 << 30), ("copies_of_2.9MB", 51680 * 56), ("copies_of_32MiB", 32 << 20)):
    n = (2 << 30) // chunk
    for rep in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i in range(n):
            dst[i * chunk:(i + 1) * chunk].copy_(big[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    out[name] = {"GB/s": n * chunk / dt / 1e9, "copies": n}
s2 = torch.cuda.Stream()
# two streams concurrently
half = 1 << 30
torch.cuda.synchronize()
t = time.perf_counter()
dst[:half].copy_(big[:half], non_blocking=True)
with torch.cuda.stream(s2):
    dst[half:].copy_(big[half:], non_blocking=True)
torch.cuda.synchronize()
out["two_streams"] = {"GB/s": (2 << 30) / (time.perf_counter() - t) / 1e9}
# file-sized copies alternating over two streams (does the second stream hide the per-copy gap?)
chunk = 51680 * 56
n = (2 << 30) // chunk
s3 = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(s2 if i & 1 else s3):
            dst[i * chunk:(i + 1) * chunk].copy_(big[i * chunk:(i + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
out["copies_of_2.9MB_two_streams"] = {"GB/s": n * chunk / dt / 1e9}
# d2h
torch.cuda.synchronize()
t = time.perf_counter()
big.copy_(dst, non_blocking=True)
torch.cuda.synchronize()
out["d2h_2GiB"] = {"GB/s": (2 << 30) / (time.perf_counter() - t) / 1e9}
print(json.dumps(out))
