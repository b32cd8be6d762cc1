#!/bin/bash
# Variants of the fused kernel's candidate walk on the bench workload: flattened pruned list (default), pruned rows, exhaustive rows through TMA staging.
run() { name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step_ms', round(d['ms_per_step'],3), 'cost', d['gn_cost_first_last'][1])"
}
run "flat minb6" PVB_FLAT=1
run "flat minb5" PVB_FLAT=1 PVB_MINB=5
run "flat minb4" PVB_FLAT=1 PVB_MINB=4
run "rows minb6" PVB_FLAT=0
run "flat minb6 again" PVB_FLAT=1
