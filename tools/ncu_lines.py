"""Attribute the warp-stall samples of an `ncu --set full --import-source on` report to SOURCE LINES.
usage: ncu_lines.py report.ncu-rep lib.so mangled_kernel_name source_file [top]
ncu's CSV source page carries SASS addresses only; nvdisasm -g of the cubin inside the library gives the address -> line map
(needs -lineinfo at compile time).  No GPU needed."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib, kern, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith("//---") and ".text." + kern + " " in l)
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith("//---"))
cur, amap = None, {}
for l in sass[start:end]:
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if mm:
        amap[int(mm.group(1), 16)] = (cur, mm.group(2))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.split("\n")))
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[h], [r for r in rows[h + 1:] if len(r) == len(rows[h])]
ix = {n: i for i, n in enumerate(hdr)}
base = int(data[0][0], 16)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
by = defaultdict(lambda: [0, defaultdict(int), 0])
for r in data:
    cur = amap.get(int(r[0], 16) - base, (None, ""))[0]
    n = int(r[ix["# Samples"]])
    by[cur][0] += n
    by[cur][2] += int(r[ix["Instructions Executed"]])
    for s in stalls:
        v = int(r[ix[s]])
        if v:
            by[cur][1][s] += v
if os.environ.get("NCU_LINES_RANGE"):   # stall histogram of a line range of the source file: NCU_LINES_RANGE=474-545
    lo, hi = map(int, os.environ["NCU_LINES_RANGE"].split("-"))
    hist, n, ins = defaultdict(int), 0, 0
    for k, v in by.items():
        if k and k[0] == os.path.basename(srcfile) and lo <= k[1] <= hi:
            n += v[0]
            ins += v[2]
            for a, b in v[1].items():
                hist[a] += b
    print(f"lines {lo}-{hi}: samples {n}, instructions {ins}")
    print(sorted(hist.items(), key=lambda x: -x[1]))
src = {}
tot = sum(v[0] for v in by.values())
print(f"total samples {tot}")
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{a[6:]} {b}" for a, b in sorted(v[1].items(), key=lambda x: -x[1])[:3])
    text = ""
    if k:
        cand = [os.path.join(os.path.dirname(srcfile), k[0]), srcfile]
        for c in cand:
            if os.path.exists(c) and os.path.basename(c) == k[0]:
                src.setdefault(c, open(c).read().split("\n"))
                text = src[c][k[1] - 1].strip()[:100]
                break
    print(f"{v[0]:6d} {100 * v[0] / tot:5.1f}%  inst {v[2]:9d}  {k}  [{st}] | {text}")
