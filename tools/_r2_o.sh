mkdir -p gpurun_out
L=gpurun_out/r2o_sweep_tight.log
: > $L
for t in 0.8 0.6 0.45; do echo "PVB_TIGHT=$t" >> $L; PVB_TIGHT=$t SWEEP_VARIANTS=4:1 timeout 600 python tools/sweep_modes.py >> $L 2>&1; done
cat $L
