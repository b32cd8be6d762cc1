mkdir -p gpurun_out
SWEEP_VARIANTS=1:0,2:1,4:1,4:0 timeout 900 python tools/sweep_modes.py > gpurun_out/r2u_sweep.log 2>&1; cat gpurun_out/r2u_sweep.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu --tb=short 2>&1 | tail -3
