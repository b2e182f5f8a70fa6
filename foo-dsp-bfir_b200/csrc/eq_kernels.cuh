// Equalizer coefficient generator on the device: the reference's
// equalizer::render_f / render_d (brutefir/equalizer.cpp:212-299, 307-394) -- 31 ISO bands + DC + Nyquist,
// raised-cosine interpolation of magnitude and phase per bin, linear-phase centring, one HC2R transform
// of `taps` points, upper half = the taps/2-sample filter -- plus the large complex transform it needs
// (taps reaches 262144 for BASELINE configs[2], beyond one CTA): a four-step decomposition
// M = M1 * M2 over two launches of the CTA-resident FFT with strided access.
#pragma once
#include "fft_core.cuh"

namespace bfir {

#define BFIR_EQ_BANDS 33 // BAND_COUNT + 2, equalizer.hpp:13,61-63

struct EqArgs {
    double freq[BFIR_EQ_BANDS];   // already divided by the sampling rate (equalizer.cpp:118)
    double mag[BFIR_EQ_BANDS];    // linear (pow(10, dB/20), :119)
    double phase[BFIR_EQ_BANDS];  // "/ (180 * M_PI)" like the reference (:120)
    void *zout;                   // [taps/2] complex: packed spectrum Z' whose inverse transform is the real filter
    const void *tw;               // exp(-2 pi i j / taps)
    int taps;
};

struct CfftArgs {
    const void *in;
    void *out;
    long long in_stride, in_batch, out_stride, out_batch; // elements (complex)
    const void *tw;              // exp(-2 pi i j / NTW)
    int tw_shift_sub;            // log2(NTW / sub-transform size)
    int apply_tw;                // multiply result k of batch b by W_Mtot^(+-k b)
    int tw_shift_tot;            // log2(NTW / Mtot)
    int mtot_mask;               // Mtot - 1
};

#ifdef __CUDACC__
// equalizer::cosine_int_f / _d (equalizer.cpp:183-204). In the float build the differences are formed
// in float and everything else in double (the 0.5 and M_PI literals promote), then narrowed on return.
template <class T> __device__ __forceinline__ T eq_cosine_int(T mag1, T mag2, T freq1, T freq2, T curfreq)
{
    const double a = __dmul_rn((double)(T)(mag1 - mag2), 0.5);
    const double x = __dmul_rn(3.14159265358979323846, (double)(T)(curfreq - freq1)) / (double)(T)(freq2 - freq1);
    const double b = __dmul_rn((double)(T)(mag1 + mag2), 0.5);
    return (T)__dadd_rn(__dmul_rn(a, cos(x)), b);
}

// X_n of the half-complex spectrum the reference writes into rbuf (equalizer.cpp:237-259 / 332-354)
template <class T> __device__ __forceinline__ cpx<T> eq_bin(const EqArgs &a, const T *eqmag, const T *eqfreq, const T *eqphase, int n)
{
    const int taps = a.taps;
    const T scale = (T)(1.0 / (T)taps), divtaps = (T)(1.0 / (T)taps);
    if (n == 0) return mk<T>(eqmag[0] * scale, (T)0);
    if (n == (taps >> 1)) return mk<T>(eqmag[BFIR_EQ_BANDS - 1] * scale, (T)0);
    const T tapspi = (T)(-(double)(T)taps * 3.14159265358979323846);
    const T curfreq = (T)n * divtaps;
    int i = 0;
    while (i < BFIR_EQ_BANDS - 2 && curfreq > eqfreq[i + 1]) i++;
    const T mag = eq_cosine_int<T>(eqmag[i], eqmag[i + 1], eqfreq[i], eqfreq[i + 1], curfreq) * scale;
    const T ph = eq_cosine_int<T>(eqphase[i], eqphase[i + 1], eqfreq[i], eqfreq[i + 1], curfreq);
    T rad, c, s;
    if (sizeof(T) == 4) {
        rad = (T)__fadd_rn(__fmul_rn((float)tapspi, (float)curfreq), (float)ph);
        c = (T)cosf((float)rad); s = (T)sinf((float)rad);
        return mk<T>((T)__fmul_rn((float)c, (float)mag), (T)__fmul_rn((float)s, (float)mag));
    }
    rad = (T)__dadd_rn(__dmul_rn((double)tapspi, (double)curfreq), (double)ph);
    c = (T)cos((double)rad); s = (T)sin((double)rad);
    return mk<T>((T)__dmul_rn((double)c, (double)mag), (T)__dmul_rn((double)s, (double)mag));
}

// one thread per k in [0, taps/2): Z'_k = (X_k + conj X_{M-k}) + i conj(W_N^k) (X_k - conj X_{M-k})
template <class T>
__global__ void eq_spectrum_kernel(const EqArgs a)
{
    __shared__ T eqmag[BFIR_EQ_BANDS], eqfreq[BFIR_EQ_BANDS], eqphase[BFIR_EQ_BANDS];
    if (threadIdx.x < BFIR_EQ_BANDS) {     // equalizer.cpp:226-231: the tables are narrowed to T first
        eqmag[threadIdx.x] = (T)a.mag[threadIdx.x];
        eqfreq[threadIdx.x] = (T)a.freq[threadIdx.x];
        eqphase[threadIdx.x] = (T)a.phase[threadIdx.x];
    }
    __syncthreads();
    const int M = a.taps >> 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const cpx<T> xk = eq_bin<T>(a, eqmag, eqfreq, eqphase, k);
    const cpx<T> xm = eq_bin<T>(a, eqmag, eqfreq, eqphase, M - k);
    const T er = xk.x + xm.x, ei = xk.y - xm.y;
    const T dr = xk.x - xm.x, di = xk.y + xm.y;
    const cpx<T> w = ((const cpx<T> *)a.tw)[k];
    const T pr = w.x * dr + w.y * di, pi = w.x * di - w.y * dr;
    ((cpx<T> *)a.zout)[k] = mk<T>(er - pi, ei + pr);
}

// batched complex transform with strided access; grid.x = batch
template <class T, int LOG2M, bool INV>
__global__ void __launch_bounds__((1 << LOG2M) / 16) cfft_strided_kernel(const CfftArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx<T> *smem = reinterpret_cast<cpx<T> *>(smem_raw);
    constexpr int NT = (1 << LOG2M) / 16;
    const int t = threadIdx.x;
    const long long b = blockIdx.x;
    const cpx<T> *in = (const cpx<T> *)a.in + b * a.in_batch;
    cpx<T> v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = in[(long long)(t + i * NT) * a.in_stride];
    fft_passes<T, LOG2M, INV, 0, 0>::run(t, v, smem, (const cpx<T> *)a.tw, a.tw_shift_sub);
    cpx<T> *out = (cpx<T> *)a.out + b * a.out_batch;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int k = t + i * NT;
        cpx<T> r = v[i];
        if (a.apply_tw) {
            const int idx = (int)(((long long)k * b) & a.mtot_mask);
            cpx<T> w = ((const cpx<T> *)a.tw)[(long long)idx << a.tw_shift_tot];
            if (INV) w = cconj(w);
            r = cmul(r, w);
        }
        out[(long long)k * a.out_stride] = r;
    }
}
#endif

} // namespace bfir
