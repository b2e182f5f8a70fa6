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
    # AddressSanitizer on the host build: the emulated kernels index exactly-sized buffers, so an
    # out-of-bounds shared / global access of the device code is caught here, without a GPU
    # (-DEMU_QUICK: sizes up to 2^10 points per CTA; the full list is for manual runs and takes minutes to compile)
    r = subprocess.run(["nvcc", "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                        "-fsanitize=address -fno-omit-frame-pointer", "-Xlinker", "-lasan", "-DEMU_QUICK", "-o", exe, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0")
    r = subprocess.run([exe, "quick"], capture_output=True, text=True, env=env)
    assert "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:]
