"""Turn the ncu outputs of one bench command into the committed summaries under profiles/:
  ncu_summaries.py launches <launches.csv> <out.json>     per-kernel totals and shares of a launch list
  ncu_summaries.py kernel <report.ncu-rep> <out.json> <offsets_per_launch>   key counters of one --set full capture"""
import csv
import json
import re
import subprocess
import sys
from collections import defaultdict

mode = sys.argv[1]
if mode == "launches":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("sgz::", "")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        tot[name][0] += 1
        tot[name][1] += ms
    allms = sum(v[1] for v in tot.values())
    ks = [{"kernel": k, "launches": v[0], "total_ms": round(v[1], 4), "share_of_all": round(v[1] / allms, 4)}
          for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])]
    search = {k: v[1] for k, v in tot.items() if k.startswith(("k_corr", "k_replay", "k_candidates", "k_refine", "k_filemax"))}
    s = sum(search.values())
    json.dump({"source": sys.argv[2], "note": "per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes",
               "kernels": ks, "search_step_shares": {k: round(v / s, 4) for k, v in sorted(search.items(), key=lambda kv: -kv[1])}},
              open(sys.argv[3], "w"), indent=1)
else:
    txt = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.split("\n")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
            "sm__pipe_tensor_cycles_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct",
            "sm__inst_executed_pipe_fp64.avg.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "smsp__cycles_active.avg", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "sm__warps_active.avg.pct",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg ", "smsp__inst_executed.sum "]
    out = {}
    for i, h in enumerate(hdr):
        if any(h.startswith(w.strip()) or (w.strip() in h) for w in want):
            out[h] = {"value": vals[i], "unit": units[i]}
    def num(name):
        for i, h in enumerate(hdr):
            if h == name:
                v = float(vals[i].replace(",", ""))
                u = units[i]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
        return None
    n_off = int(sys.argv[4])
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    json.dump({"source": sys.argv[2], "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "k_corr_tc2",
               "offsets_per_launch": n_off, "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
               "dram_bytes_per_offset": (rd + wr) / n_off, "algorithmic_bytes_per_offset": 64,
               "counters": out}, open(sys.argv[3], "w"), indent=1)
