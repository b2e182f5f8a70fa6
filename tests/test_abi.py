"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses loudly to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, load_package

HEADER = os.path.join(ROOT, "include", "bfir_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bfir_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_reference_entry_points():
    syms = header_symbols()
    # the reference methods the north star names (fftw_convolver.hpp / brutefir.hpp)
    for name in ["convolve", "convolve_add", "convolve_inplace", "crossfade_inplace", "mixnscale", "time2freq",
                 "freq2time", "dirac_convolve", "dirac_convolve_inplace", "convolve_eval", "raw2cbuf", "cbuf2raw",
                 "coeffs2cbuf", "runtime_coeffs2cbuf", "cbufsize"]:
        assert "bfir_conv_" + name in syms
    for name in ["create", "destroy", "is_initialized", "set_coeff", "run", "reset", "check_overflows"]:
        assert "bfir_" + name in syms


def test_library_exports_every_declared_symbol():
    pkg = load_package()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_table_matches_header():
    pkg = load_package()
    bound = sorted(name for name, _, _ in pkg.API)
    assert bound == header_symbols()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    pkg = load_package()
    with pytest.raises(pkg.BfirError) as e:
        pkg.Brutefir(64, 2, 4, 2, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    assert e.value.code == pkg.ERR_CUDA
    with pytest.raises(pkg.BfirError):
        pkg.FftwConvolver(64, 4)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under the package may reference it"""
    pdir = os.path.join(ROOT, "foo-dsp-bfir_b200")
    for dirpath, _, files in os.walk(pdir):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "libbfir_ref", "libbfir_oracle", "oracle/_", "fft_r2r"):
                    assert needle not in text, (dirpath, f, needle)


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/bfir_b200.h must compile as C99 (a cgo / JNI / P-Invoke binding parses it as C)"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not found")
    src = tmp_path / "h.c"
    src.write_text('#include "bfir_b200.h"\nint main(void) { bfir_config_t c; (void)c; return BFIR_OK; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
