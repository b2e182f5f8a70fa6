// Stand-alone kernels behind the reference's per-function convolver entry points (the "inner"
// surface, brutefir/fftw_convolver.hpp:39-133) and the serial dither requantiser. They operate on
// device cbufs in the reference's own layouts, so each can be compared with the oracle on identical
// buffers. Arithmetic uses the reference's operation order with explicitly unfused multiplies and
// adds (the reference is an SSE2 build without FMA), which makes these entry points bit-exact.
#pragma once
#include "rfft_kernels.cuh"

namespace bfir {

enum { MIXMODE_INPUT = 1, MIXMODE_INPUT_ADD = 2, MIXMODE_OUTPUT = 3 }; // fftw_convolver.hpp:14-16
#define BFIR_MAX_MIX_BUFS 64

struct MixArgs {
    const void *in[BFIR_MAX_MIX_BUFS];
    double scales[BFIR_MAX_MIX_BUFS];
    void *out;
    int n_bufs, mixmode, N;
};

struct DitherState {          // device image of dither_state_t (global.h:63-69)
    int randtab_ptr;
    int tab0;                 // private copy of dither_randtab[0] (the reference shares one byte, dither.cpp:133)
    double err[2];            // sf[] / sd[] error feedback, stored widened
};

struct DitherArgs {
    const void *real;         // [channels][L] reals
    void *raw;                // interleaved raw output
    long long raw_stream_stride; // bytes
    int fmt, ch_per_stream, L, n_channels;
    const int8_t *randtab;
    int randtab_size;
    const void *randmap;      // 512 entries [-256..255] of T
    DitherState *dstate;      // [channels]
    OverflowStats *stats;     // [channels]
    int single_channel;       // >= 0: inner API call for exactly this dither channel (grid of 1)
    long long real_stride;    // elements between channels of `real`
    int ch_base;              // first channel of this launch; n_channels counts from it
};

#ifdef __CUDACC__
template <class T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <class T> __device__ __forceinline__ T sub_rn(T a, T b) { return add_rn<T>(a, -b); }

// convolver_mixnscale, fftw_convolver.cpp:859-1427 (float) / :1559-2123 (double). One thread per ORD slot.
template <class T>
__global__ void mixnscale_kernel(const MixArgs a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x; // index in the ORD buffer
    if (j >= a.N) return;
    const int N = a.N, half = N >> 1;
    const int g = j >> 3, lane = j & 7;
    int k = 4 * g + (lane & 3);
    int hc; // matching index in the HC buffer
    if (lane < 4) hc = k;
    else hc = (k == 0) ? half : N - k;
    const int src = a.mixmode == MIXMODE_INPUT ? hc : j;
    const int dst = a.mixmode == MIXMODE_INPUT ? j : hc;
    T acc = mul_rn<T>(((const T *)a.in[0])[src], (T)a.scales[0]);
    for (int b = 1; b < a.n_bufs; b++) acc = add_rn<T>(acc, mul_rn<T>(((const T *)a.in[b])[src], (T)a.scales[b]));
    // all reads of this thread's sources happen before its write; callers keep out distinct from in
    ((T *)a.out)[dst] = acc;
}

// MODE 0: convolve (out = in*c), 1: convolve_add (out += in*c). One thread per group of 8.
// fftw_convolver.cpp:1430-1525 / :2126-2220
template <class T, int MODE>
__global__ void convolve_kernel(const T *b, const T *c, T *d, int N)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g * 8 >= N) return;
    T bb[8], cc[8], dd[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { bb[j] = b[g * 8 + j]; cc[j] = c[g * 8 + j]; dd[j] = MODE == 1 ? d[g * 8 + j] : (T)0; }
    T o[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const T re = sub_rn<T>(mul_rn<T>(bb[j], cc[j]), mul_rn<T>(bb[j + 4], cc[j + 4]));
        const T im = add_rn<T>(mul_rn<T>(bb[j], cc[j + 4]), mul_rn<T>(bb[j + 4], cc[j]));
        o[j] = MODE == 1 ? add_rn<T>(dd[j], re) : re;
        o[j + 4] = MODE == 1 ? add_rn<T>(dd[j + 4], im) : im;
    }
    if (g == 0) {
        const T d1 = mul_rn<T>(bb[0], cc[0]), d2 = mul_rn<T>(bb[4], cc[4]);
        o[0] = MODE == 1 ? add_rn<T>(dd[0], d1) : d1;
        o[4] = MODE == 1 ? add_rn<T>(dd[4], d2) : d2;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) d[g * 8 + j] = o[j];
}

// convolver_dirac_convolve[_inplace], fftw_convolver.cpp:1528-1556 / :2223-2251 (HC layout)
template <class T>
__global__ void dirac_kernel(const T *in, T *out, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const T fraction = (T)1.0 / (T)N;
    out[i] = mul_rn<T>(in[i], (i & 1) ? -fraction : fraction);
}

// convolve_inplace_ordered, fftw_convolver.cpp:820-856: complex product on the plain HC layout
// (Re at n, Im at size - n); one thread per bin, both halves of a bin are read before either is written
template <class T>
__global__ void hc_convolve_inplace_kernel(T *b, const T *c, int size)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x, size2 = size >> 1;
    if (n > size2) return;
    if (n == 0 || n == size2) { b[n] = mul_rn<T>(b[n], c[n]); return; }
    const T br = b[n], bi = b[size - n], cr = c[n], ci = c[size - n];
    b[n] = sub_rn<T>(mul_rn<T>(br, cr), mul_rn<T>(bi, ci));
    b[size - n] = add_rn<T>(mul_rn<T>(br, ci), mul_rn<T>(bi, cr));
}

// linear old->new ramp of convolver_crossfade_inplace, float branch fftw_convolver.cpp:296-305, used
// for both precisions (the double branch :306-315 reads memory the function never wrote)
template <class T>
__device__ __forceinline__ T crossfade_ramp(T old_v, T new_v, int n, int L)
{
    if (sizeof(T) == 4) {
        const float f = (float)(1.0 / (double)(float)(L - 1));
        const float fn = __fmul_rn(f, (float)n);
        const double a = __dmul_rn((double)old_v, __dadd_rn(1.0, -(double)fn));
        const float b = __fmul_rn(__fmul_rn((float)new_v, f), (float)n);
        return (T)(float)__dadd_rn(a, (double)b);
    } else {
        const double d = 1.0 / (double)(L - 1);
        const double dn = __dmul_rn(d, (double)n);
        const double a = __dmul_rn((double)old_v, __dadd_rn(1.0, -dn));
        const double b = __dmul_rn(__dmul_rn((double)new_v, d), (double)n);
        return (T)__dadd_rn(a, b);
    }
}

template <class T>
__global__ void crossfade_ramp_kernel(const T *xfade, T *buffer, int L)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= L) return;
    buffer[n] = crossfade_ramp<T>(xfade[n], buffer[n], n, L);
}

// Engine-side filter swap (BASELINE configs[2]): the block is computed with the old and the new
// coefficient set, both accumulated spectra are brought to the time domain, and the first L samples are
// the linear old->new ramp of convolver_crossfade_inplace (fftw_convolver.cpp:296-305). This kernel is
// the ramp fused with the output stage (probe + cbuf2raw); the re-FFT of crossfade_inplace (:317-320)
// is skipped because the consumer is the output stage.
struct XfadeArgs {
    const void *t_old, *t_new;   // [channels][N] time-domain outputs of the old / new filter
    void *out;                   // raw interleaved output, or planar reals [channels][L] for the dither kernel
    long long out_stream_stride; // bytes
    int N, L, fmt, ch_per_stream, ch_base, to_real;
    double ovf_max;
    OverflowStats *stats;
    EngineState *state;
    int *host_flag;
};

template <class T>
__global__ void __launch_bounds__(256) xfade_emit_kernel(const XfadeArgs a)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y + a.ch_base;
    OverflowAcc acc;
    acc.n_overflows = 0; acc.intlargest = 0; acc.largest = 0.0;
    if (n < a.L) {
        const T y = crossfade_ramp<T>(((const T *)a.t_old)[(long long)ch * a.N + n], ((const T *)a.t_new)[(long long)ch * a.N + n], n, a.L);
        if (n == 0 && !(y - y == (T)0)) {                                               // brutefir.cpp:316-321
            atomicMin(&a.state->first_bad_channel, ch);
            if (a.host_flag != NULL) *(volatile int *)a.host_flag = 1;
        }
        if (a.to_real) {
            ((T *)a.out)[(long long)ch * a.L + n] = y;
        } else {
            const int stream = ch / a.ch_per_stream, c = ch - stream * a.ch_per_stream;
            const int bytes = fmt_bytes(a.fmt);
            uint8_t *p = (uint8_t *)a.out + (long long)stream * a.out_stream_stride + ((long long)n * a.ch_per_stream + c) * bytes;
            if (fmt_isfloat(a.fmt)) store_raw_float<T>(p, a.fmt, y, (T)a.ovf_max, acc);
            else {
                int32_t imin, imax;
                int_limits(a.fmt, imin, imax);
                store_raw_quantised<T>(p, a.fmt, y, (T)imin, (T)imax, imin, imax, acc);
            }
        }
    }
    if (!a.to_real) overflow_commit(&a.stats[ch], acc);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.state->blockcounter += 1u;
}

// convolver_raw2cbuf, fftw_convolver.cpp:157-185: L strided raw samples -> next_cbuf[0..L) and cbuf[L..2L)
template <class T>
__global__ void raw2cbuf_kernel(const uint8_t *raw, T *cbuf, T *next_cbuf, int fmt, int spacing, int L)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= L) return;
    const T x = load_raw<T>(raw + (long long)n * spacing * fmt_bytes(fmt), fmt);
    next_cbuf[n] = x;
    cbuf[L + n] = x;
}

// real2raw*_no_dither on a planar real buffer (inner API cbuf2raw without dither), one thread per sample
template <class T>
__global__ void real2raw_kernel(const T *real, uint8_t *raw, int fmt, int spacing, int L, double ovf_max, OverflowStats *stats)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    OverflowAcc acc;
    acc.n_overflows = 0; acc.intlargest = 0; acc.largest = 0.0;
    if (n < L) {
        uint8_t *p = raw + (long long)n * spacing * fmt_bytes(fmt);
        if (fmt_isfloat(fmt)) {
            store_raw_float<T>(p, fmt, real[n], (T)ovf_max, acc);
        } else {
            int32_t imin, imax;
            int_limits(fmt, imin, imax);
            store_raw_quantised<T>(p, fmt, real[n], (T)imin, (T)imax, imin, imax, acc);
        }
    }
    if (acc.n_overflows) atomicAdd(&stats->n_overflows, acc.n_overflows);
    if (acc.intlargest > 0) atomicMax(&stats->intlargest, acc.intlargest);
    if (acc.largest > 0.0) atomicMax(&stats->largest_bits, (unsigned long long)__double_as_longlong(acc.largest));
}

// quantise() of codec.cuh for the dither walker, without a branch: every branch of the original sits on the walker's
// dependent chain (one warp, nothing else to issue). The value the error feedback subtracts, (T)s, is computed in
// floating point -- trunc, minus one for negative values, clamped to the format's range: the correctly rounded image
// of the same integer, so it equals (T)s bit for bit (also where T cannot hold s exactly: 32-bit samples in single
// precision) -- while the integer s itself, which only the output needs, is formed beside the chain with selects,
// like the statistics. Same results as quantise() for every finite input.
template <class T>
__device__ __forceinline__ T quantise_walk(T d, T stat, T rmin, T rmax, int32_t imin, int32_t imax, int32_t &s_out, OverflowAcc &acc)
{
    const bool neg = d < (T)0;
    const T tr = sizeof(T) == 4 ? (T)truncf((float)d) : (T)trunc((double)d);
    const T sf = fmin(fmax(add_rn<T>(tr, neg ? (T)-1 : (T)0), rmin), rmax);   // + 0 also turns -0 into the +0 of (T)int
    const bool ovn = neg && d <= rmin, ovp = !neg && d > rmax;
    int32_t s = real_to_int_rz<T>(d) - (neg ? 1 : 0);
    s = ovn ? imin : (ovp ? imax : s);
    acc.n_overflows += (ovn || ovp) ? 1u : 0u;
    const double ds = (double)stat, dd = (double)d;
    const bool ln = ovn && ds < -acc.largest, lp = ovp && ds > acc.largest;
    acc.largest = ln ? -dd : (lp ? dd : acc.largest);
    const bool in = neg && !ovn && s < -acc.intlargest, ip = !neg && !ovp && s > acc.intlargest;
    acc.intlargest = in ? -s : (ip ? s : acc.intlargest);
    s_out = s;
    return sf;
}

// real2raw*_hp_tpdf -> dither*_real2int_hp_tpdf (real2raw.cpp:39-315 / :645-920, dither.cpp:127-212 /
// :276-347): first-order high-pass error feedback + TPDF dither. The recurrence
//     x[n] = real[n] + (e[n-1] - e[n-2]);  s[n] = Q(x[n] + dv[n]);  e[n] = x[n] - s[n]
// is serial in n (Q rounds and clamps), so ONE lane walks a channel -- but nothing else has to be serial. A CTA of
// four warps takes G channels (G = 1 .. 32, a power of two chosen by the launch so that few-channel engines get
// one channel per CTA and many-channel ones fill the lanes of the walker warp) and moves through the block in
// chunks of DITHER_STAGE / G samples, double buffered in shared memory:
//     warps 1-3: stage chunk k+1 -- real[] (coalesced) and the dither values dv[n] = map[tab[base+n] - tab[base+n-1]],
//                which do not depend on the recurrence -- and write chunk k-1's samples out in the raw format
//                (lanes along the interleaved channels)
//     warp 0:    lane g walks channel g through chunk k reading and writing shared memory only
// so the walker's dependent chain is sub, add, add, trunc, add, min, max, sub (quantise_walk: no branch, no integer
// round trip) and no global-memory latency sits on it
// (round 1: one thread per channel with every operand fetched from global memory inside the loop, 230 ns per sample).
#define DITHER_STAGE 1024
template <class T>
__global__ void __launch_bounds__(128) dither_kernel(const DitherArgs a, const int G)
{
    constexpr int PAD = DITHER_STAGE + DITHER_STAGE / 32;       // odd channel pitch G | 1 needs up to 1/32 more
    __shared__ T real_s[2][PAD], dv_s[2][PAD];
    __shared__ int32_t out_s[2][PAD];
    __shared__ int base_s[32], tab0_s[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g0 = blockIdx.x * G;                              // first channel of this CTA, relative to the launch
    const int ng = min(G, a.n_channels - g0);
    const int GP = G | 1;                                       // shared-memory pitch between consecutive samples
    const int CH = DITHER_STAGE / G;                            // samples per chunk and channel
    const int nchunks = (a.L + CH - 1) / CH;
    const T *map = (const T *)a.randmap + 256;
    // channel g of this CTA: state / statistics index, data index (planar reals, raw position)
    const int single = a.single_channel;
    // walker registers (warp 0, lane < ng)
    DitherState st = {};
    OverflowAcc acc;
    acc.n_overflows = 0; acc.intlargest = 0; acc.largest = 0.0;
    T e0 = (T)0, e1 = (T)0;
    const int my_ch = single >= 0 ? single : a.ch_base + g0 + lane;
    if (warp == 0 && lane < ng) {
        st = a.dstate[my_ch];
        // dither_preloop_real2int_hp_tpdf, dither.cpp:127-139
        if (st.randtab_ptr + a.L >= a.randtab_size) {
            st.tab0 = a.randtab[st.randtab_ptr - 1];
            st.randtab_ptr = 1;
        }
        base_s[lane] = st.randtab_ptr;
        tab0_s[lane] = st.tab0;
        st.randtab_ptr += a.L;
        const OverflowStats os = a.stats[my_ch];     // sequential, so seed with the running values (dither.cpp:170-205)
        acc.intlargest = os.intlargest;
        acc.largest = __longlong_as_double((long long)os.largest_bits);
        e0 = (T)st.err[0]; e1 = (T)st.err[1];
    }
    __syncthreads();

    auto stage = [&](int k, int first, int nthreads) {
        const int n0 = k * CH, len = min(CH, a.L - n0);
        T *rs = real_s[k & 1], *ds = dv_s[k & 1];
        for (int idx = first; idx < ng * len; idx += nthreads) {
            const int g = idx / len, n = idx - g * len;
            const int data_ch = single >= 0 ? 0 : a.ch_base + g0 + g;
            const int ti = base_s[g] + n0 + n;
            const int cur = (int)a.randtab[ti];
            const int prev = ti - 1 == 0 ? tab0_s[g] : (int)a.randtab[ti - 1];
            rs[n * GP + g] = ((const T *)a.real)[(long long)data_ch * a.real_stride + n0 + n];
            ds[n * GP + g] = map[cur - prev];
        }
    };
    auto emit = [&](int k, int first, int nthreads) {
        const int n0 = k * CH, len = min(CH, a.L - n0);
        const int32_t *os = out_s[k & 1];
        const int bytes = fmt_bytes(a.fmt);
        for (int idx = first; idx < ng * len; idx += nthreads) {
            const int n = idx / ng, g = idx - n * ng;
            const int data_ch = single >= 0 ? 0 : a.ch_base + g0 + g;
            const int stream = data_ch / a.ch_per_stream, c = data_ch - stream * a.ch_per_stream;
            uint8_t *p = (uint8_t *)a.raw + (long long)stream * a.raw_stream_stride + ((long long)(n0 + n) * a.ch_per_stream + c) * bytes;
            store_raw_int(p, a.fmt, os[n * GP + g]);
        }
    };

    int32_t imin, imax;
    int_limits(a.fmt, imin, imax);
    const T rmin = (T)imin, rmax = (T)imax;
    stage(0, tid, 128);
    __syncthreads();
    for (int k = 0; k < nchunks; k++) {
        if (warp == 0) {
            if (lane < ng) {
                const int len = min(CH, a.L - k * CH);
                const T *rs = real_s[k & 1] + lane, *ds = dv_s[k & 1] + lane;
                int32_t *os = out_s[k & 1] + lane;
                for (int n = 0; n < len; n += 8) {              // L and CH are powers of two >= 16
                    T r[8], dv[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) { r[j] = rs[(n + j) * GP]; dv[j] = ds[(n + j) * GP]; }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const T x = add_rn<T>(r[j], sub_rn<T>(e0, e1));      // error feedback {1, -1}
                        e1 = e0;
                        const T d = add_rn<T>(x, dv[j]);
                        int32_t s;
                        const T sf = quantise_walk<T>(d, x, rmin, rmax, imin, imax, s, acc);
                        e0 = sub_rn<T>(x, sf);
                        os[(n + j) * GP] = s;
                    }
                }
            }
        } else {
            if (k + 1 < nchunks) stage(k + 1, tid - 32, 96);
            if (k > 0) emit(k - 1, tid - 32, 96);
        }
        __syncthreads();
    }
    emit(nchunks - 1, tid, 128);
    if (warp == 0 && lane < ng) {
        st.err[0] = (double)e0;
        st.err[1] = (double)e1;
        a.dstate[my_ch] = st;
        OverflowStats *os = &a.stats[my_ch];
        os->n_overflows += acc.n_overflows;
        os->intlargest = acc.intlargest;
        os->largest_bits = (unsigned long long)__double_as_longlong(acc.largest);
    }
}

// channels per CTA of the dither kernel: one while that still gives at most two CTAs per SM's worth of walkers,
// then powers of two up to a full walker warp
static inline int dither_channels_per_cta(int n_channels)
{
    int g = 1;
    while (g < 32 && (n_channels + g - 1) / g > 296) g *= 2;
    return g;
}

// engine housekeeping
#define BFIR_MAX_GROUPS 8
static __global__ void engine_reset_kernel(EngineState *state, int *procblocks, OverflowStats *stats, int n_filters, int n_stats)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < BFIR_MAX_GROUPS) { state[i].blockcounter = 0; state[i].first_bad_channel = 0x7fffffff; }
    // brutefir.cpp:347-367: counters only, never the buffers
    if (i < n_filters) procblocks[i] = 0;
    if (i < n_stats) { stats[i].n_overflows = 0; stats[i].intlargest = 0; stats[i].largest_bits = 0ull; }
}

// undo the bookkeeping of a block that brutefir::run would have aborted at channel `bad`
// (brutefir.cpp:316-321 returns before :337-340): later channels never ran, the counter did not move
static __global__ void engine_abort_fixup_kernel(EngineState *state, int n_groups, int *procblocks, const unsigned char *pb_inc, int n_channels, int bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups) { state[i].blockcounter -= 1u; state[i].first_bad_channel = 0x7fffffff; }
    if (i > bad && i < n_channels && pb_inc[i]) procblocks[i] -= 1;
}
#endif

} // namespace bfir
