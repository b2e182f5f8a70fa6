// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle, never linked into the product library.
//
// Self-contained power-of-two real FFT in the two flavours the reference asks FFTW for:
// FFTW_R2HC (real -> half-complex, unnormalised) and FFTW_HC2R (half-complex -> real,
// unnormalised), cf. the plan creation at reference brutefir/fftw_convolver.cpp:796-807.
//
// FFTW 3.3-beta1 itself (fftw/libfftw3-3.dll, PE32 binary only) cannot be used on this host, so the
// published definition of the two transforms is restated here:
//     R2HC:  hc[k]   = Re X_k (0<=k<=N/2),  hc[N-k] = Im X_k (0<k<N/2),  X_k = sum_n x_n e^{-2 pi i k n/N}
//     HC2R:  x_n = sum_{k=0}^{N-1} X_k e^{+2 pi i k n/N}  (X_{N-k} = conj X_k), i.e. N * inverse
// Arithmetic is done in the native precision T (float for fftwf_*, double for fftw_*), like FFTW.
// Algorithm: N-point real transform through an N/2-point complex Stockham radix-4 FFT.
#pragma once
#include <cmath>
#include <complex>
#include <vector>
#include <cstddef>

namespace oracle_fft {

template <class T>
class RealFFT {
public:
    typedef std::complex<T> cpx;

    explicit RealFFT(int n_real) : N(n_real), M(n_real / 2) {
        // twiddles for the N/2-point complex FFT, one table per stage
        int n = M;
        while (n > 1) {
            int radix = (n % 4 == 0) ? 4 : 2;
            int m = n / radix;
            std::vector<cpx> t((size_t)m * (radix - 1));
            for (int p = 0; p < m; p++)
                for (int k = 1; k < radix; k++) {
                    long double a = -2.0L * M_PIl * (long double)p * k / n;
                    t[(size_t)p * (radix - 1) + (k - 1)] = cpx((T)cosl(a), (T)sinl(a));
                }
            stage_tw.push_back(t);
            stage_radix.push_back(radix);
            n = m;
        }
        split_tw.resize(M + 1);
        for (int k = 0; k <= M; k++) {
            long double a = -2.0L * M_PIl * (long double)k / N;
            split_tw[k] = cpx((T)cosl(a), (T)sinl(a));
        }
        bufa.resize(M > 0 ? M : 1);
        bufb.resize(M > 0 ? M : 1);
    }

    int size() const { return N; }

    // in: N reals, out: N reals in FFTW half-complex order. in == out allowed.
    void r2hc(const T *in, T *out) {
        if (N == 1) { out[0] = in[0]; return; }
        for (int n = 0; n < M; n++) bufa[n] = cpx(in[2 * n], in[2 * n + 1]);
        cpx *Z = cfft<false>(bufa.data(), bufb.data());
        T zr = Z[0].real(), zi = Z[0].imag();
        out[0] = zr + zi;
        out[M] = zr - zi;
        for (int k = 1; k < M; k++) {
            cpx a = Z[k], b = std::conj(Z[M - k]);
            cpx e = (a + b) * (T)0.5;
            cpx d = (a - b) * (T)0.5;
            cpx o(d.imag(), -d.real()); // d / i
            cpx x = e + split_tw[k] * o;
            out[k] = x.real();
            out[N - k] = x.imag();
        }
    }

    // in: N reals in half-complex order, out: N reals (unnormalised). in == out allowed.
    void hc2r(const T *in, T *out) {
        if (N == 1) { out[0] = in[0]; return; }
        T x0 = in[0], xm = in[M];
        bufa[0] = cpx(x0 + xm, x0 - xm);
        for (int k = 1; k < M; k++) {
            // X_k for 0<k<M is (in[k], in[N-k]); X_{M-k} likewise
            cpx xk(in[k], in[N - k]);
            int j = M - k;
            cpx xj(in[j], in[N - j]);
            cpx b = std::conj(xj);
            cpx e = xk + b;
            cpx d = (xk - b) * std::conj(split_tw[k]);
            bufa[k] = e + cpx(-d.imag(), d.real()); // e + i d
        }
        cpx *z = cfft<true>(bufa.data(), bufb.data());
        for (int n = 0; n < M; n++) {
            out[2 * n] = z[n].real();
            out[2 * n + 1] = z[n].imag();
        }
    }

private:
    // Stockham autosort; returns the buffer that holds the result
    template <bool INV>
    cpx *cfft(cpx *x, cpx *y) {
        int n = M, s = 1;
        for (size_t st = 0; st < stage_radix.size(); st++) {
            const int radix = stage_radix[st];
            const cpx *tw = stage_tw[st].data();
            const int m = n / radix;
            if (radix == 4) {
                for (int p = 0; p < m; p++) {
                    cpx w1 = tw[3 * p], w2 = tw[3 * p + 1], w3 = tw[3 * p + 2];
                    if (INV) { w1 = std::conj(w1); w2 = std::conj(w2); w3 = std::conj(w3); }
                    const cpx *x0 = x + (size_t)s * p;
                    const cpx *x1 = x0 + (size_t)s * m;
                    const cpx *x2 = x1 + (size_t)s * m;
                    const cpx *x3 = x2 + (size_t)s * m;
                    cpx *y0 = y + (size_t)s * 4 * p;
                    for (int q = 0; q < s; q++) {
                        T ar = x0[q].real(), ai = x0[q].imag();
                        T br = x1[q].real(), bi = x1[q].imag();
                        T cr = x2[q].real(), ci = x2[q].imag();
                        T dr = x3[q].real(), di = x3[q].imag();
                        T apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
                        T bpdr = br + dr, bpdi = bi + di, bmdr = br - dr, bmdi = bi - di;
                        // forward: -i*(b-d) = (bmdi, -bmdr); inverse: +i*(b-d) = (-bmdi, bmdr)
                        T jr = INV ? -bmdi : bmdi, ji = INV ? bmdr : -bmdr;
                        T t1r = amcr + jr, t1i = amci + ji;
                        T t2r = apcr - bpdr, t2i = apci - bpdi;
                        T t3r = amcr - jr, t3i = amci - ji;
                        y0[q] = cpx(apcr + bpdr, apci + bpdi);
                        y0[q + s] = cpx(t1r * w1.real() - t1i * w1.imag(), t1r * w1.imag() + t1i * w1.real());
                        y0[q + 2 * s] = cpx(t2r * w2.real() - t2i * w2.imag(), t2r * w2.imag() + t2i * w2.real());
                        y0[q + 3 * s] = cpx(t3r * w3.real() - t3i * w3.imag(), t3r * w3.imag() + t3i * w3.real());
                    }
                }
            } else {
                for (int p = 0; p < m; p++) {
                    cpx w1 = tw[p];
                    if (INV) w1 = std::conj(w1);
                    const cpx *x0 = x + (size_t)s * p;
                    const cpx *x1 = x0 + (size_t)s * m;
                    cpx *y0 = y + (size_t)s * 2 * p;
                    for (int q = 0; q < s; q++) {
                        cpx a = x0[q], b = x1[q];
                        cpx d = a - b;
                        y0[q] = a + b;
                        y0[q + s] = cpx(d.real() * w1.real() - d.imag() * w1.imag(),
                                        d.real() * w1.imag() + d.imag() * w1.real());
                    }
                }
            }
            cpx *t = x; x = y; y = t;
            n = m;
            s *= radix;
        }
        return x;
    }

    int N, M;
    std::vector<std::vector<cpx> > stage_tw;
    std::vector<int> stage_radix;
    std::vector<cpx> split_tw;
    std::vector<cpx> bufa, bufb;
};

} // namespace oracle_fft
