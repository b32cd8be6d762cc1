mkdir -p gpurun_out
L=gpurun_out/r2h_sweep.log
: > $L
for ro in 0.5 1.0 2.0; do echo "PVB_REORDER=$ro" >> $L; PVB_REORDER=$ro SWEEP_VARIANTS=4:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1; done
echo "PVB_STATIC=0" >> $L; PVB_STATIC=0 SWEEP_VARIANTS=4:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1
echo "hints off (static bound only) / mode 2" >> $L; SWEEP_VARIANTS=4:0,2:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1
cat $L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_zz_gpu_reference_fixtures.py -x -q -m gpu --tb=short 2>&1 | tail -8
