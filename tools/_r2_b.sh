mkdir -p gpurun_out
SWEEP_VARIANTS=1:0,2:1,3:1 timeout 900 python tools/sweep_modes.py > gpurun_out/r2b_sweep_modes.log 2>&1
cat gpurun_out/r2b_sweep_modes.log
