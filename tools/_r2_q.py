import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panovlm_b200
from tools import bench_configs
ctx = panovlm_b200.Context(0)
print(json.dumps(bench_configs.floor_refine_pose(ctx, 1, 0)))
