"""Key counters of the latency-bound tail kernels from their `ncu --set full` captures (tools/evidence_run.sh):
python tools/ncu_tail.py out.json report1.ncu-rep report2.ncu-rep ...   (no GPU needed)"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]
out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.split("\n")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"source": rep}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            rec["kernel"] = v
        if h in WANT:
            rec[h] = {"value": v, "unit": u}
    stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
              for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    rec["warp_stalls_per_issue_active (top 5)"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    out[rec.get("kernel", rep)] = rec
json.dump(out, open(sys.argv[1], "w"), indent=1)
