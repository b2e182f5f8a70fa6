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


TOL = {4: 1e-5, 8: 1e-12}     # BASELINE.json north_star: relative RMS vs the reference path, float / double builds
_REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def parity(name, got, ref, tol, truth=None):
    """The parity gate of the GPU tests, with the measured figure on record.

    Passes when rel_rms(got, ref) <= tol (the north star's 1e-5 / 1e-12 against the reference path). Where two
    single-precision implementations of a chain of several transforms differ by more than that, the caller supplies
    `truth` (the same chain evaluated in float64): the gate is then "no less accurate than the reference itself",
    rel_rms(got, truth) <= max(tol, rel_rms(ref, truth)), and the record shows all three figures. Every call
    appends one JSON line to gpurun_out/parity_report.jsonl (when that directory exists) and prints it."""
    import json
    rec = {"test": name, "rel_rms_vs_oracle": rel_rms(got, ref), "tol": tol}
    if truth is not None:
        rec["gpu_vs_truth"] = rel_rms(got, truth)
        rec["oracle_vs_truth"] = rel_rms(ref, truth)
    rec["gate"] = "vs_oracle" if rec["rel_rms_vs_oracle"] <= tol else ("vs_truth" if truth is not None else "FAILED")
    print("parity:", json.dumps(rec))
    if os.path.isdir(os.path.dirname(_REPORT)):
        with open(_REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    if rec["rel_rms_vs_oracle"] <= tol:
        return rec
    assert truth is not None, "%s: rel RMS %.3e above %.1e" % (name, rec["rel_rms_vs_oracle"], tol)
    assert rec["gpu_vs_truth"] <= max(tol, rec["oracle_vs_truth"]), \
        "%s: %.3e vs the oracle, %.3e vs float64 truth (the oracle itself: %.3e)" % (
            name, rec["rel_rms_vs_oracle"], rec["gpu_vs_truth"], rec["oracle_vs_truth"])
    return rec


class SwapChain:
    """The reference's entry points composed in run() order (brutefir.cpp:245-343) with
    convolver_crossfade_inplace (fftw_convolver.cpp:276-321) between the partition sums and the output stage:
    the oracle of a filter swap on a block (BASELINE configs[2]). `cv` is an oracle.Convolver; one instance in
    the engine's precision is the oracle, a second one in double on the same inputs is the float64 truth."""

    def __init__(self, cv, L, P, C, dtype):
        self.cv, self.L, self.P, self.C, self.dt = cv, L, P, C, dtype
        self.fdl = np.zeros((C, P, 2 * L), dtype=dtype)
        self.prev = np.zeros((C, L), dtype=dtype)
        self.t = 0

    def spectra(self, h, same_for_all=False):
        """coeff::preprocess_coeff (coeff.cpp:293-354) per channel; h: list of coefficient arrays"""
        H = [self.cv.preprocess_coeff(np.asarray(f, dtype=self.dt), self.P) for f in (h[:1] if same_for_all else h)]
        return [H[0]] * self.C if same_for_all else H

    def _psum(self, c, Hc):
        t, P, cv = self.t, self.P, self.cv
        acc = cv.convolve(self.fdl[c, t % P].copy(), Hc[0].copy())
        for i in range(1, min(P, t + 1)):
            cv.convolve_add(self.fdl[c, (t - i) % P].copy(), Hc[i].copy(), acc)
        return acc

    def block(self, blk, H_new, H_old=None):
        """blk: [L, C] samples; H_old given = this block swaps H_old -> H_new with a crossfade. Returns [L, C]."""
        cv, L = self.cv, self.L
        blk = np.asarray(blk, dtype=self.dt)
        y = np.empty((L, self.C), dtype=self.dt)
        for c in range(self.C):
            self.fdl[c, self.t % self.P] = cv.mixnscale([cv.time2freq(np.concatenate([self.prev[c], blk[:, c]]))], [1.0], 1)
            self.prev[c] = blk[:, c]
            if H_old is None:
                spec = self._psum(c, H_new[c])
            else:
                spec = cv.crossfade_inplace(self._psum(c, H_new[c]), self._psum(c, H_old[c]), cv.cbuf())
            y[:, c] = cv.freq2time(cv.mixnscale([spec], [1.0], 3))[:L]
        self.t += 1
        return y


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
