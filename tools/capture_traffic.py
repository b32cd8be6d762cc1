"""Regenerates profiles/traffic.json from a LIVE ncu capture of the bench's fused kernel (run on the GPU box through gpurun):

  gpurun -- 'python tools/capture_traffic.py gpurun_out/r2_traffic'      # writes <prefix>.ncu-rep, <prefix>.txt (summary) and <prefix>.json
  cp gpurun_out/r2_traffic.json profiles/traffic.json                     # here, after the call (plus the .txt summary under profiles/)

The json carries the sha256 of the kernel sources it was captured from; bench.py ignores a traffic.json whose hash differs from the tree's
(roofline.traffic = null), so the figure on a bench line is always traceable to a capture of that very kernel.
The capture is one steady-state launch of `python bench.py` (64 frames, 10 M queries): the 8 chunked launches of the first evaluation and the
first evaluations after the warm-up poses are skipped."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "traffic")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 16            # 8 chunk launches + 2 warm-up + 6 timed evaluations
os.makedirs(os.path.dirname(prefix), exist_ok=True)
cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:k_associate", "-s", str(skip), "-c", "1", "-f", "-o", prefix,
       sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--no-extra"]
subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", prefix + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
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
       "kernel_sources_sha256_16": bench.kernel_sources_hash(), "capture": os.path.basename(prefix) + ".ncu-rep (ncu --set full, one steady-state launch of python bench.py)"}
json.dump(out, open(prefix + ".json", "w"), indent=1)
with open(prefix + ".txt", "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarize.py"), prefix + ".ncu-rep"], capture_output=True, text=True).stdout)
print(json.dumps(out))
