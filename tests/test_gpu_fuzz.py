"""Randomised parity sweep (tests/fuzz_parity_tool.py) as a GPU test: random scenes / placements / cell sizes / thresholds / k,
with and without TMA staging, association sets identical to the oracle and reduced systems equal to 1e-7."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_parity_dense_path():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity_tool.py"), "16"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["associations_checked"] > 5000 and res["worst_system_rel"] < 1e-7 and res["worst_plane_abs"] < 1e-8
