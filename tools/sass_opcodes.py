"""Per-kernel SASS opcode histogram of libsgz_b200.so (cuobjdump -sass): the mnemonics that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, SYNCS = mbarrier, USETMAXREG = setmaxnreg) next to the
arithmetic ones.  usage: sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.json   (no GPU needed)"""
import json
import os
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "strugatzki_b200", "libsgz_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
watch = ["UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "FFMA2", "FFMA", "DFMA", "DADD", "DMUL", "MUFU", "F2F",
         "LDG", "STG", "LDS", "STS", "REDG", "ATOMG", "STL", "LDL", "HMMA"]
out, cur = {}, None
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = out.setdefault(name, Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur["instructions"] += 1
        op = m.group(1)
        for w in watch:
            if op == w:
                cur[w] += 1
res = {k: {w: v[w] for w in ["instructions"] + watch if v[w]} for k, v in sorted(out.items())}
tot = Counter()
for v in res.values():
    tot.update(v)
print(json.dumps({"library": os.path.basename(lib), "source": "cuobjdump -sass", "total": dict(tot), "kernels": res}, indent=1))
