"""Regenerates profiles/traffic.json from a LIVE ncu capture of the bench's fused kernel (run on the GPU box through gpurun):

  gpurun -- 'python tools/capture_traffic.py gpurun_out/r2_traffic'      # writes <prefix>.ncu-rep, <prefix>.txt (summary) and <prefix>.json
  cp gpurun_out/r2_traffic.json profiles/traffic.json                     # here, after the call (plus the .txt summary under profiles/)

The json carries the sha256 of the kernel sources it was captured from; bench.py ignores a traffic.json whose hash differs from the tree's
(roofline.traffic = null), so the figure on a bench line is always traceable to a capture of that very kernel.
The capture is one full-grid launch of the bench workload (64 frames, 10 M queries against the 10 M-point target) in the steady state of a Gauss-Newton
run (tools/profile_dense.py brackets it with cudaProfilerStart / Stop); pass `cold` or `moved` as second argument for the other states."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "traffic")
state = sys.argv[2] if len(sys.argv) > 2 else "steady"          # steady | cold | moved (tools/profile_dense.py)
os.makedirs(os.path.dirname(prefix), exist_ok=True)
rep_prefix = os.path.join("/tmp", os.path.basename(prefix))      # the report itself (tens of MB with sources) stays out of gpurun_out
# ONE full-grid launch of the bench workload (64 frames, 10 M queries against the 10 M-point target) in the chosen state, bracketed by cudaProfilerStart/Stop
cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "--profile-from-start", "off", "-k", "regex:k_associate", "-f", "-o", rep_prefix,
       sys.executable, os.path.join(ROOT, "tools", "profile_dense.py"), state]
subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", rep_prefix + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
m = dict(zip(rows[0], zip(rows[1], rows[2])))


def val(key):
    unit, v = m[key]
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    return v * scale


import bench  # noqa: E402
out = {"kernel": m["Kernel Name"][1], "k_associate_dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
       "k_associate_warp_instructions_per_launch": val("smsp__inst_executed.sum"), "duration_ms_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")) * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(m["gpu__time_duration.sum"][0], 1.0),
       "kernel_sources_sha256_16": bench.kernel_sources_hash(), "capture": os.path.basename(prefix) + " (ncu --set full, one full-grid " + state + "-state launch of the bench workload, tools/profile_dense.py)"}
json.dump(out, open(prefix + ".json", "w"), indent=1)
with open(prefix + ".txt", "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarize.py"), rep_prefix + ".ncu-rep"], capture_output=True, text=True).stdout)
print(json.dumps(out))

sym = "_ZN3pvb11k_associateILi10ELb1ELi6ELb0ELi4ELb1EEEvNS_9AssocArgsE"
with open(prefix + "_lines.txt", "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep_prefix + ".ncu-rep", sym, "50"], capture_output=True, text=True).stdout)
