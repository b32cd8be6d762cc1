mkdir -p gpurun_out
L=gpurun_out/r2c_sweep_superrow.log
: > $L
for hs in 0.65 0.75 0.85 1.0; do echo "PVB_DENSE_HSCALE=$hs" >> $L; PVB_DENSE_HSCALE=$hs SWEEP_VARIANTS=4:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1; done
echo "mode 2 / mode 1 for the bit-identity check (first variant = mode 1)" >> $L
SWEEP_VARIANTS=1:0,2:1,4:1,4:0 timeout 900 python tools/sweep_modes.py >> $L 2>&1
cat $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_zz_gpu_reference_fixtures.py -x -q -m gpu --tb=short 2>&1 | tail -8
