"""TEST INFRASTRUCTURE ONLY -- numpy float64 restatement of the reference's convolver layouts and block
loop, independent of oracle/fft_r2r.hpp (it uses numpy.fft), used to pin the two compiled oracles and
as the "truth" for error measurements (direct linear convolution).

Layouts (reference brutefir/fftw_convolver.cpp):
  T   time cbuf   [previous block | current block]                       :168-184
  HC  FFTW half-complex  hc[k]=Re X_k (0<=k<=N/2), hc[N-k]=Im X_k        :798-806
  ORD convolver order: groups of 8 = [Re k..k+3 | Im k..k+3]; slot 4 = Re X_{N/2}   :884-905
"""
import numpy as np


def r2hc(x):
    """FFTW_R2HC, unnormalised."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    X = np.fft.rfft(x)
    hc = np.empty(n)
    hc[: n // 2 + 1] = X.real
    hc[n // 2 + 1:] = X.imag[1: n // 2][::-1]
    return hc


def hc2r(hc):
    """FFTW_HC2R, unnormalised (= N * inverse)."""
    hc = np.asarray(hc, dtype=np.float64)
    n = len(hc)
    X = np.zeros(n // 2 + 1, dtype=np.complex128)
    X.real = hc[: n // 2 + 1]
    X.imag[1: n // 2] = hc[n // 2 + 1:][::-1]
    return np.fft.irfft(X, n) * n


def hc_to_ord(hc, scale=1.0):
    """mixnscale MIXMODE_INPUT with one buffer (fftw_convolver.cpp:883-907)."""
    hc = np.asarray(hc, dtype=np.float64)
    n = len(hc)
    half = n // 2
    re = hc[:half].copy()                       # Re X_0 .. Re X_{half-1}
    im = np.empty(half)
    im[0] = hc[half]                            # Nyquist takes the place of Im X_0
    im[1:] = hc[n - 1: half: -1]                # Im X_1 .. Im X_{half-1}
    out = np.empty(n)
    o = out.reshape(-1, 8)
    o[:, :4] = re.reshape(-1, 4)
    o[:, 4:] = im.reshape(-1, 4)
    return out * scale


def ord_to_hc(o, scale=1.0):
    """mixnscale MIXMODE_OUTPUT with one buffer (fftw_convolver.cpp:1163-1186)."""
    o = np.asarray(o, dtype=np.float64)
    n = len(o)
    half = n // 2
    g = o.reshape(-1, 8)
    re = g[:, :4].reshape(-1)
    im = g[:, 4:].reshape(-1)
    hc = np.empty(n)
    hc[:half] = re
    hc[half] = im[0]
    hc[n - 1: half: -1] = im[1:]
    return hc * scale


def ord_to_complex(o):
    """ORD -> complex spectrum X_0..X_{N/2}."""
    o = np.asarray(o, dtype=np.float64)
    half = len(o) // 2
    g = o.reshape(-1, 8)
    re = g[:, :4].reshape(-1)
    im = g[:, 4:].reshape(-1).copy()
    ny = im[0]
    im[0] = 0.0
    return np.concatenate([re + 1j * im, [ny + 0j]])


def complex_to_ord(X):
    half = len(X) - 1
    re = X[:half].real.copy()
    im = X[:half].imag.copy()
    im[0] = X[half].real
    out = np.empty(2 * half)
    g = out.reshape(-1, 8)
    g[:, :4] = re.reshape(-1, 4)
    g[:, 4:] = im.reshape(-1, 4)
    return out


def convolve_ord(b, c):
    """convolver_convolve on ORD buffers (fftw_convolver.cpp:1465-1493)."""
    return complex_to_ord(ord_to_complex(b) * ord_to_complex(c))


def coeffs2cbuf(coeffs, L, scale=1.0):
    """fftw_convolver.cpp:475-537: partition in the UPPER half, R2HC, HC->ORD / N."""
    n = 2 * L
    r = np.zeros(n)
    k = min(len(coeffs), L)
    r[L: L + k] = np.asarray(coeffs[:k], dtype=np.float64) * scale
    return hc_to_ord(r2hc(r), 1.0 / n)


def preprocess_coeff(coeffs, L, blocks, scale=1.0):
    """coeff.cpp:293-354."""
    coeffs = np.asarray(coeffs, dtype=np.float64)
    out = np.zeros((blocks, 2 * L))
    for i in range(blocks):
        out[i] = coeffs2cbuf(coeffs[i * L: (i + 1) * L], L, scale)
    return out


class EngineNP:
    """brutefir::run (brutefir.cpp:245-343) for planar float64 input/output, no quantisation."""

    def __init__(self, L, P, C):
        self.L, self.N, self.P, self.C = L, 2 * L, P, C
        self.fdl = np.zeros((C, P, self.N))
        self.prev = np.zeros((C, L))
        self.coeffs = None
        self.procblocks = 0
        self.blockcounter = 0

    def set_coeff(self, coeffs, blocks, scale=1.0):
        self.coeffs = np.stack([preprocess_coeff(c, self.L, blocks, scale) for c in coeffs])
        self.coeff_blocks = blocks

    def run(self, x, in_scale=1.0, out_scale=1.0):
        """x: [C, L] planar -> y: [C, L] planar."""
        y = np.empty((self.C, self.L))
        self.procblocks = min(self.procblocks + 1, self.P)
        cur = self.blockcounter % self.P
        for n in range(self.C):
            t = np.concatenate([self.prev[n], x[n]])
            self.prev[n] = x[n]
            self.fdl[n, cur] = hc_to_ord(r2hc(t), in_scale)
            acc = ord_to_complex(self.fdl[n, cur]) * ord_to_complex(self.coeffs[n, 0])
            for i in range(1, min(self.coeff_blocks, self.procblocks)):
                acc = acc + ord_to_complex(self.fdl[n, (self.blockcounter - i) % self.P]) * ord_to_complex(self.coeffs[n, i])
            y[n] = hc2r(ord_to_hc(complex_to_ord(acc), out_scale))[: self.L]
        self.blockcounter += 1
        return y


def direct_convolution(x, h, n_out):
    """First n_out samples of the linear convolution x * h (the truth run() must reproduce)."""
    from scipy.signal import fftconvolve
    return fftconvolve(np.asarray(x, dtype=np.float64), np.asarray(h, dtype=np.float64))[:n_out]


def rel_rms(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = np.sqrt(np.mean(b * b))
    return float(np.sqrt(np.mean((a - b) ** 2)) / den) if den > 0 else float(np.sqrt(np.mean((a - b) ** 2)))
