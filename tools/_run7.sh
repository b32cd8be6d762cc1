for b in 12 14 15; do echo "MORTON_BITS=$b"; PVB_MORTON_BITS=$b SWEEP_VARIANTS=2:1,3:1 python tools/sweep_modes.py 2>&1 | tail -2; done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batched or estimate_pose or odometry" 2>&1 | tail -5
