for f in 0 1; do echo "FLAT=$f"; PVB_FLAT=$f SWEEP_VARIANTS=2:1 python tools/sweep_modes.py 2>&1 | tail -1; done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_reduced_block.py tests/test_zz_gpu_reference_fixtures.py -x -q -m gpu --tb=short 2>&1 | tail -15
