"""Row f4 on the device: tracer advection with a prescribed mass flux against the oracle (tests/gpu_tracer_parity.py).
First run on hardware: the driver's round-1 GPU suite (GPUTEST_r01.json: XPASS on all four cases), so the test is a plain
parity test now; it still runs in a subprocess so that a device fault stays contained."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tracer_advection_prescribed_mass_flux():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_tracer_parity.py")], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:]); print(out.stderr[-2000:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
