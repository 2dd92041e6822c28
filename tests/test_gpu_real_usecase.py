"""The device path on the reference's own test catchment, from the committed fixture (tests/golden/realcase_*.npz): static
state from the real input maps, raw float32 meteo maps through the feeder kernel, soil, overland and split channel routing;
against the CPU restatement's recorded run and against the soil-moisture maps of the output stacks the reference SHIPS
(tools/real_usecase_device_check.py, run as its own process).

This check was written after the GPU budget of the round was spent and has NOT been run on a GPU yet; the catchment has
properties no other GPU test has (every pixel is a channel pixel, real parameter ranges).  It is therefore marked as a
non-strict expected failure: XPASS = the device reproduces the shipped outputs, XFAIL = a discrepancy to look at."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="never run on a GPU yet (added after the round's GPU budget was spent)")
def test_device_reproduces_the_shipped_soil_moisture(gpu_lib):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "real_usecase_device_check.py")], cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-4000:])
    assert "REAL USECASE PASSED" in r.stdout, r.stdout[-3000:]
