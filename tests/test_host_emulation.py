"""CPU test of the CUDA FFT kernels' logic: tests/host_emulation/emu_fft.cu executes every phase of
rfft_forward_kernel / rfft_inverse_kernel for all threads on the host (same __host__ __device__ code
the GPU runs) and compares with the oracle FFT. Needs nvcc (host compilation only), no GPU."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not found")
def test_block_fft_host_emulation(tmp_path):
    exe = str(tmp_path / "emu_fft")
    src = os.path.join(ROOT, "tests", "host_emulation", "emu_fft.cu")
    r = subprocess.run(["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:]
