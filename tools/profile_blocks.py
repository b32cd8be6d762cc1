"""Runs the Room-shaped K3/K4 measurement of bench.py standalone (for ncu captures of k_eval_blocks)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import panovlm_b200  # noqa: E402

ctx = panovlm_b200.Context(0)
print(json.dumps(bench.bench_blocks(ctx, 6490.2)))
ctx.close()
