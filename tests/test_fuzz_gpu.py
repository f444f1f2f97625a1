"""Short randomised differential runs (a few seconds each) of the projection and isotonic-regression kernels against
the oracle, bit-exact: random layouts, inputs and call configurations (tools/pava_fuzz.py, tools/proj_fuzz.py), and of the
three device-resident solver loops against the oracle's BATCH restatement (tools/solver_fuzz.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("tool,seed0", [("pava_fuzz.py", 100000), ("proj_fuzz.py", 100000), ("solver_fuzz.py", 100000)])
def test_randomised_differential(tool, seed0):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), "6", str(seed0)], capture_output=True, text=True, timeout=300,
                         cwd=ROOT)
    assert res.returncode == 0, res.stdout[-500:] + res.stderr[-2000:]
    assert res.stdout.strip().startswith("ok ")
