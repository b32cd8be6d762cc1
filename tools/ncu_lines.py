"""Source-line view of an `ncu --set full --import-source on` capture without the GUI (development tool):
joins the per-SASS-instruction metrics of `ncu --page source --csv` with the line table of the cubin inside the shipped library
(`nvdisasm -gi`) and prints the executed warp instructions / stall samples per source line, per inlined call chain root and per
source function region.  The library on disk must be the one the capture ran (same SASS).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep '_ZN3pvb11k_associateILi10ELb1ELi6ELb0ELi2ELb1EEEvNS_9AssocArgsE' [top_n]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, sym = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "panovlm_b200", "libpanovlm_b200.so")], cwd=tmp, capture_output=True)
cubin = None
for f in os.listdir(tmp):
    if f.endswith(".cubin") and sym.encode() in open(os.path.join(tmp, f), "rb").read():
        cubin = os.path.join(tmp, f)
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
lines, chain, on, chain_open = {}, [], False, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = ln.startswith(".text." + sym + ":")
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if not chain_open:
            chain, chain_open = [], True
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        chain_open = False
        lines[int(m.group(1), 16)] = (list(chain), m.group(2).strip())
    else:
        chain_open = False

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ai, si, ii, ti, smp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
base = None
per_line, per_root, per_second, per_third = (defaultdict(lambda: [0, 0, 0]) for _ in range(4))

tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    try:
        addr, n, t, s = int(r[ai], 16), int(r[ii]), int(r[ti]), int(r[smp])
    except ValueError:
        continue
    if base is None:
        base = addr
    off = addr - base
    ch, _ = lines.get(off, ([("?", 0)], ""))
    inner = ch[0] if ch else ("?", 0)
    root = ch[-1] if ch else ("?", 0)
    second = ch[-2] if len(ch) >= 2 else root
    third = ch[-3] if len(ch) >= 3 else second
    for d, k in ((per_line, inner), (per_root, root), (per_second, second), (per_third, third)):
        d[k][0] += n; d[k][1] += t; d[k][2] += s
    tot_i += n; tot_s += s


def show(title, d):
    print(f"\n== {title} (share of executed warp instructions | share of stall samples | avg active lanes)")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print(f"{100 * v[0] / tot_i:5.1f}%  {100 * v[2] / max(1, tot_s):5.1f}%  {v[1] / max(1, v[0]):4.1f}   {k[0]}:{k[1]}")


print(f"total executed warp instructions {tot_i}, stall samples {tot_s}")
show("by kernel-level line (root of the inline chain)", per_root)
show("by the line one level below the kernel (e.g. the statement of associate_point2plane)", per_second)
show("two levels below the kernel (e.g. the statement of knn_select_hinted)", per_third)
show("by innermost source line", per_line)
