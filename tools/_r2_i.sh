mkdir -p gpurun_out
L=gpurun_out/r2i_sweep.log
: > $L
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pchk tools/micro/packed_sqdist_check.cu && /tmp/pchk >> $L 2>&1
SWEEP_VARIANTS=4:1,4:0 timeout 600 python tools/sweep_modes.py >> $L 2>&1
echo "PVB_STATIC=0" >> $L; PVB_STATIC=0 SWEEP_VARIANTS=4:1,4:0 timeout 600 python tools/sweep_modes.py >> $L 2>&1
cat $L
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_zz_gpu_reference_fixtures.py -x -q -m gpu --tb=short 2>&1 | tail -8
