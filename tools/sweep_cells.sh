#!/bin/bash
# Sweep of the pruned block walk: start radius (PVB_R0) x cell size on the bench workload; prints kernel ms per variant.
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --no-e2e $CELLARG 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step_ms', round(d['ms_per_step'],3), 'cost', d['gn_cost_first_last'][1])"
}
CELLARG="" run "r0=1 h=auto(0.159)" PVB_R0=1
CELLARG="--cell 0.12" run "r0=1 h=0.12" PVB_R0=1 PVB_CELLCAP=8
CELLARG="--cell 0.12" run "r0=2 h=0.12" PVB_R0=2 PVB_CELLCAP=8
CELLARG="--cell 0.10" run "r0=2 h=0.10" PVB_R0=2 PVB_CELLCAP=8
CELLARG="--cell 0.08" run "r0=2 h=0.08" PVB_R0=2 PVB_CELLCAP=16
CELLARG="--cell 0.065" run "r0=2 h=0.065" PVB_R0=2 PVB_CELLCAP=32
