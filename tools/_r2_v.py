import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panovlm_b200
from tools import bench_configs
ctx = panovlm_b200.Context(0)
r = bench_configs.room_joint(ctx)
print(json.dumps(r)[:1500])
