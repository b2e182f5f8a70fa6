"""pytest configuration: `gpu` marker, package / oracle loaders, synthetic signal helpers."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_package():
    return importlib.import_module("foo-dsp-bfir_b200")


@pytest.fixture(scope="session")
def pkg():
    """The product package; the shared library must already be built (python __graft_entry__.py)."""
    p = load_package()
    p.load_library()
    return p


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle front-end (TEST INFRASTRUCTURE): builds the port on demand, uses oracle/_ref if present."""
    import oracle as o
    if not o.available("port"):
        o.build("port")
    if not o.available("ref") and os.path.isdir("/root/reference/brutefir"):
        o.build("ref")
    return o


def oracle_kinds():
    import oracle as o
    kinds = ["port"]
    if o.available("ref") or os.path.isdir("/root/reference/brutefir"):
        kinds.append("ref")
    return kinds


# ---------------------------------------------------------------- synthetic data
def white_noise(seed, n_frames, n_channels):
    """uniform white noise in [-1, 1), like buffer::load_white_noise (reference buffer.cpp:455-493) but seeded"""
    return np.random.default_rng(0xB200 + seed).uniform(-1.0, 1.0, size=(n_frames, n_channels))


def decay_filter(ch, taps):
    """exponentially decaying Gaussian filter, unit L2 norm"""
    g = np.random.default_rng(1000 + ch).standard_normal(taps)
    h = g * np.exp(-6.9 * np.arange(taps) / taps)
    return h / np.sqrt(np.sum(h * h))


def rel_rms(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.sqrt(np.mean(b * b))
    err = np.sqrt(np.mean((a - b) ** 2))
    return float(err / den) if den > 0 else float(err)


FMT_NP = {1: np.int8, 2: "<i2", 3: ">i2", 6: "<i4", 7: ">i4", 8: "<f4", 9: ">f4", 10: "<f8", 11: ">f8"}


def encode_raw(x, fmt):
    """float frames [n, C] in [-1, 1) -> interleaved raw bytes in sample format `fmt`"""
    x = np.asarray(x, dtype=np.float64)
    if fmt in (8, 9, 10, 11):
        return np.ascontiguousarray(x.astype(FMT_NP[fmt])).view(np.uint8).ravel()
    bits = {1: 8, 2: 16, 3: 16, 4: 24, 5: 24, 6: 32, 7: 32}[fmt]
    v = np.clip(np.round(x * (2 ** (bits - 1))), -(2 ** (bits - 1)), 2 ** (bits - 1) - 1).astype(np.int64)
    if bits != 24:
        return np.ascontiguousarray(v.astype(FMT_NP[fmt])).view(np.uint8).ravel()
    u = (v & 0xFFFFFF).astype(np.uint32).ravel()
    b = np.stack([(u >> 0) & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=1).astype(np.uint8)
    if fmt == 5:
        b = b[:, ::-1]
    return np.ascontiguousarray(b).ravel()


def decode_raw(raw, fmt, n_channels):
    """interleaved raw bytes -> [n, C] array of the sample values (ints for integer formats)"""
    raw = np.asarray(raw, dtype=np.uint8)
    if fmt not in (4, 5):
        return raw.view(FMT_NP[fmt]).reshape(-1, n_channels).astype(np.float64 if fmt >= 8 else np.int64)
    b = raw.reshape(-1, 3).astype(np.int64)
    if fmt == 5:
        b = b[:, ::-1]
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    v = np.where(v >= 1 << 23, v - (1 << 24), v)
    return v.reshape(-1, n_channels)
