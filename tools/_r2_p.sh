mkdir -p gpurun_out
L=gpurun_out/r2p_sweep_reorder.log
: > $L
for ro in 0.25 0.5 1.0; do echo "PVB_REORDER=$ro" >> $L; PVB_REORDER=$ro SWEEP_VARIANTS=4:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1; PVB_REORDER=$ro python bench.py --no-cpu-baseline --no-extra --no-e2e 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench value', l['value'], 'kernel_ms', l['roofline']['kernel_ms'], l['roofline']['kernel_ms_per_step'])" >> $L; done
cat $L
