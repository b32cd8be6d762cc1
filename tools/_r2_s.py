import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panovlm_b200
from tools import bench_configs
ctx = panovlm_b200.Context(0)
for rep in range(2):
    r = bench_configs.room_refine_pose(ctx)
    print(r["estimate_pose_s"], [(p["iterations"], round(p["build_s"], 3), round(p["lm_s"], 3)) for p in r["per_outer"]])
