"""Turns an .ncu-rep (ncu --set full, one k_associate launch) into the text summary committed under profiles/.
Usage: python profiles/summarize.py gpurun_out/prof_assoc_r1X.ncu-rep > profiles/r1X_k_associate.txt"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, zip(units, vals)))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}")
for k in keys:
    if k in m:
        print(f"{k:75s} {m[k][1]} {m[k][0]}")
st = sorted(((float(v[1]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h, v in m.items()
             if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
print("warp stall reasons (warps per issue): " + " | ".join(f"{h}={v:.2f}" for v, h in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
si, ii, ti = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
agg, tot = defaultdict(lambda: [0, 0]), 0
for r in rows[2:]:
    if len(r) < len(h2):
        continue
    try:
        n, t = int(r[ii]), int(r[ti])
    except ValueError:
        continue
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si].strip())
    op = mm.group(2).split(".")[0] if mm else r[si][:10]
    agg[op][0] += n; agg[op][1] += t; tot += n
print("executed warp instructions by opcode (share, avg active lanes): " +
      " | ".join(f"{k} {100 * v[0] / tot:.1f}% ({v[1] / max(1, v[0]):.0f})" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]))
