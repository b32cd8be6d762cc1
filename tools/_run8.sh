for f in 0 1; do echo "FLAT=$f"; PVB_FLAT=$f SWEEP_VARIANTS=2:1 python tools/sweep_modes.py 2>&1 | tail -1; done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=short -k "estimate_pose" 2>&1 | tail -30
