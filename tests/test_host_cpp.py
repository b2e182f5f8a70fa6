"""The C++ host mirror (foo-dsp-bfir_b200/host/*.hpp) compiles against the C ABI and behaves like the
reference classes: CPU run checks the no-device failure contract, GPU run filters audio."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "host_cpp", "host_mirror_test.cpp")
SRC_TOOLS = os.path.join(ROOT, "tests", "host_cpp", "host_tools_test.cpp")
LIBDIR = os.path.join(ROOT, "foo-dsp-bfir_b200")


def build(tmp_path, src=SRC):
    exe = str(tmp_path / os.path.basename(src).replace(".cpp", ""))
    r = subprocess.run(["g++", "-std=c++11", "-O1", src, "-o", exe, "-L" + LIBDIR, "-lbfir_b200",
                        "-Wl,-rpath," + LIBDIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not found")
def test_host_mirror_refuses_without_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = subprocess.run([build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_filters_audio(tmp_path):
    r = subprocess.run([build(tmp_path), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not found")
def test_plugin_adapter_pass_through_without_filter(tmp_path):
    """dsp_bfir::on_chunk contract (foo_dsp_bfir.cpp:352-357): no filter -> chunk passes through"""
    r = subprocess.run([build(tmp_path, SRC_TOOLS)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_plugin_adapter_and_offline_tools(tmp_path):
    """next rows N2 (dsp_impl_base framing) and N3 (convolve_impulses, calculate_attenuation) on the GPU engine"""
    r = subprocess.run([build(tmp_path, SRC_TOOLS), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_offline_tools_match_reference_preprocessor(tmp_path):
    """N3: host/preprocessor.hpp on the GPU engine against the reference's own preprocessor.cpp (oracle/_ref, fed the
    same dense responses from in-memory sound files): convolve_impulses 1e-5 / 1e-12, calculate_attenuation in dB"""
    refdir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(refdir, "libbfir_ref.so")):
        pytest.skip("oracle/_ref/libbfir_ref.so not built")
    src = os.path.join(ROOT, "tests", "host_cpp", "preprocessor_parity_test.cpp")
    exe = str(tmp_path / "preprocessor_parity_test")
    r = subprocess.run(["g++", "-std=c++11", "-O1", src, "-o", exe, "-L" + LIBDIR, "-lbfir_b200", "-L" + refdir, "-lbfir_ref",
                        "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + refdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
