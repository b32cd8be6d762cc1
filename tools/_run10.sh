SWEEP_VARIANTS=2:1 python tools/sweep_modes.py 2>&1 | tail -1
python tools/capture_traffic.py gpurun_out/r2c_traffic 2>&1 | tail -2
python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=short -k "dense or estimate" 2>&1 | tail -3
