"""Randomised schedules over every run entry point (tools/stress_async.py): bfir_run, bfir_run_device[_pipelined],
bfir_run_device_pair / _quad / _oct (joined and staged), bfir_run_async, bfir_run_async_pair, bfir_run_async_quad with
random waits / syncs / joins in between must all produce the one-block device path's output."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_schedules(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_async.py"), "8", str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "stress ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
