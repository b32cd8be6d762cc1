mkdir -p gpurun_out
L=gpurun_out/r2d_sweep_cellorder.log
: > $L
for hs in 0.85 1.0 1.15; do echo "PVB_DENSE_HSCALE=$hs" >> $L; PVB_DENSE_HSCALE=$hs SWEEP_VARIANTS=4:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1; done
SWEEP_VARIANTS=1:0,4:0 timeout 900 python tools/sweep_modes.py >> $L 2>&1
cat $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu --tb=short -k "dense or fuzz" 2>&1 | tail -8
