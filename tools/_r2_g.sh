mkdir -p gpurun_out
L=gpurun_out/r2g_sweep.log
SWEEP_VARIANTS=1:0,2:1,4:1 timeout 900 python tools/sweep_modes.py > $L 2>&1
cat $L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_zz_gpu_reference_fixtures.py -x -q -m gpu --tb=short 2>&1 | tail -8
