mkdir -p gpurun_out
L=gpurun_out/r2j_sweep.log
: > $L
echo "LC4=32 prefetch" >> $L
SWEEP_VARIANTS=4:1,4:0 timeout 600 python tools/sweep_modes.py >> $L 2>&1
echo "LC4=24 prefetch" >> $L
cp panovlm_b200/libpanovlm_b200.so /tmp/keep.so; cp gpurun_lib24.so panovlm_b200/libpanovlm_b200.so
SWEEP_VARIANTS=4:1,4:0 timeout 600 python tools/sweep_modes.py >> $L 2>&1
cp /tmp/keep.so panovlm_b200/libpanovlm_b200.so
cat $L
