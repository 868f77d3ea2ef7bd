"""Short run of the randomised parity sweep (tools/stress_parity.py) inside the GPU suite."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_randomised_parity_sweep(cuda):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_parity.py"), "--cases", "16", "--seed", "11"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "0 failures" in res.stdout
