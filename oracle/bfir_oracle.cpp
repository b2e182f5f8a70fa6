// TEST INFRASTRUCTURE ONLY -- CPU oracle ("port"). Never linked into, imported by or called from the
// product library; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the .so built from this file.
//
// A from-scratch restatement of the reference's partitioned overlap-save convolver, one function per
// reference function, each citing the file:line (relative to /root/reference) it follows. It exists
// so that (1) the GPU box, where /root/reference is absent, always has a CPU checker even if the
// prebuilt oracle/_ref/libbfir_ref.so (the unmodified reference sources) did not travel, and (2) the
// reference build can itself be cross-checked by an independent statement of the same algorithm.
// Pinning: tests/test_oracle.py checks this file against oracle/_ref (reference sources run here),
// against tests/golden/*.npz (vectors generated from oracle/_ref by tests/golden/make_golden.py) and
// against oracle/oracle_np.py (numpy float64 + direct linear convolution). The reference ships no
// golden vectors or known-answer tests of its own, and its FFT (FFTW 3.3-beta1,
// binary-only) is replaced by oracle/fft_r2r.hpp in both oracle builds.
//
// Differences from the reference, all deliberate and documented in DESIGN.md:
//   * no BF_MAXCHANNELS limit (global.h:21);
//   * crossfade uses the float-branch algorithm for double too (fftw_convolver.cpp:296-305 vs the
//     broken 306-315);
//   * dither map entry +255 is defined (0.5 + 256/255) instead of read out of bounds (dither.cpp:77-78);
//   * runtime_coeffs2cbuf uses per-instance scratch (fftw_convolver.cpp:543-549 uses a static).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "fft_r2r.hpp"

namespace {

enum { FMT_S8 = 1, FMT_S16_LE, FMT_S16_BE, FMT_S24_LE, FMT_S24_BE, FMT_S32_LE, FMT_S32_BE,
       FMT_FLOAT_LE, FMT_FLOAT_BE, FMT_FLOAT64_LE, FMT_FLOAT64_BE };
enum { MIX_INPUT = 1, MIX_INPUT_ADD = 2, MIX_OUTPUT = 3 };

struct overflow_t { // image of bfoverflow_t, global.h:96-102
    unsigned int n_overflows;
    int32_t intlargest;
    double largest;
    double max;
};

struct sample_format { // global.h:39-47 filled like brutefir.cpp:436-539
    bool isfloat, swap;
    int bytes;
    double scale;
};

int fill_format(sample_format *sf, int format, bool normalized)
{
    switch (format) {
    case FMT_S8: sf->bytes = 1; sf->isfloat = false; sf->swap = false; break;
    case FMT_S16_LE: sf->bytes = 2; sf->isfloat = false; sf->swap = false; break;
    case FMT_S16_BE: sf->bytes = 2; sf->isfloat = false; sf->swap = true; break;
    case FMT_S24_LE: sf->bytes = 3; sf->isfloat = false; sf->swap = false; break;
    case FMT_S24_BE: sf->bytes = 3; sf->isfloat = false; sf->swap = true; break;
    case FMT_S32_LE: sf->bytes = 4; sf->isfloat = false; sf->swap = false; break;
    case FMT_S32_BE: sf->bytes = 4; sf->isfloat = false; sf->swap = true; break;
    case FMT_FLOAT_LE: sf->bytes = 4; sf->isfloat = true; sf->swap = false; break;
    case FMT_FLOAT_BE: sf->bytes = 4; sf->isfloat = true; sf->swap = true; break;
    case FMT_FLOAT64_LE: sf->bytes = 8; sf->isfloat = true; sf->swap = false; break;
    case FMT_FLOAT64_BE: sf->bytes = 8; sf->isfloat = true; sf->swap = true; break;
    default: return -1;
    }
    if (sf->isfloat) sf->scale = 1.0;
    else {
        double full = (double)(1 << ((sf->bytes << 3) - 1)); // brutefir.cpp:398-414
        sf->scale = normalized ? 1.0 / full : full;
    }
    return 0;
}

// ------------------------------------------------------------------ raw2real (raw2real.cpp:17-424)
// One sample: assemble the little-endian value of `bytes` bytes (after optional byte swap) and convert.
template <class T>
void raw2real(T *real, const uint8_t *raw, const sample_format &sf, int spacing, int n_samples)
{
    const int stride = spacing * sf.bytes;
    for (int n = 0; n < n_samples; n++, raw += stride) {
        uint8_t b[8];
        for (int i = 0; i < sf.bytes; i++) b[i] = sf.swap ? raw[sf.bytes - 1 - i] : raw[i];
        if (sf.isfloat) {
            if (sf.bytes == 4) { float f; memcpy(&f, b, 4); real[n] = (T)f; }   // :42-58, :246-265
            else { double d; memcpy(&d, b, 8); real[n] = (T)d; }                 // :59-76, :266-280
        } else {
            int32_t v;
            switch (sf.bytes) {
            case 1: v = (int8_t)b[0]; break;                                      // :86-91
            case 2: v = (int16_t)((uint16_t)b[0] | ((uint16_t)b[1] << 8)); break; // :92-125
            case 3: // 3 bytes into the top of an int32, arithmetic >> 8 (:126-176)
                v = (int32_t)(((uint32_t)b[0] << 8) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 24)) >> 8;
                break;
            default: v = (int32_t)((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24)); break;
            }
            real[n] = (T)v;
        }
    }
}

// ------------------------------------------------------------------ dither (dither.cpp)
struct dither_state { // dither_state_t, global.h:63-69; `tab0` privatises dither_randtab[0], see below
    int randtab_ptr;
    int base;      // index of randtab[] that loop_counter 0 maps to (state->randtab - dither_randtab)
};

#define TAUSWORTHE(s, a, b, c, d) ((((s) & (c)) << (d)) ^ ((((s) << (a)) ^ (s)) >> (b)))

struct Dither {
    std::vector<int8_t> tab;
    std::vector<double> mapd; // randmap[-256..255] (index +256); entry 255 is the documented extension
    std::vector<float> mapf;
    std::vector<dither_state> st;
    int size, realsize;

    static uint32_t tausrand(uint32_t s[3]) // dither.cpp:418-426
    {
        s[0] = TAUSWORTHE(s[0], 13, 19, 4294967294U, 12);
        s[1] = TAUSWORTHE(s[1], 2, 25, 4294967288U, 4);
        s[2] = TAUSWORTHE(s[2], 3, 11, 4294967280U, 17);
        return s[0] ^ s[1] ^ s[2];
    }

    Dither(int n_channels, int sample_rate, int rs, int max_size, int max_samples_per_loop) // dither.cpp:21-110
    {
        realsize = rs;
        int spacing = 10 * sample_rate;
        int minspacing = (sample_rate > max_samples_per_loop) ? sample_rate : max_samples_per_loop;
        if (spacing < minspacing) spacing = minspacing;
        if (max_size > 0 && n_channels * spacing > max_size) spacing = max_size / n_channels;
        size = n_channels * spacing + 1;
        uint32_t s[3];
        uint32_t seed = 1; // tausinit(state, 0) -> default seed 1 (dither.cpp:429-449)
        s[0] = (69069u * seed) & 0xFFFFFFFFu;
        s[1] = (69069u * s[0]) & 0xFFFFFFFFu;
        s[2] = (69069u * s[1]) & 0xFFFFFFFFu;
        for (int i = 0; i < 6; i++) tausrand(s);
        tab.resize(size);
        for (int n = 0; n < size; n++) tab[n] = (int8_t)(tausrand(s) & 0xFF);
        mapd.resize(512);
        mapf.resize(512);
        mapd[0] = -0.5; mapf[0] = -0.5f;                        // [-256], dither.cpp:82,94
        for (int n = -255; n < 254; n++) {
            // float build: the expression is evaluated in double and stored to float (dither.cpp:86-87)
            mapf[n + 256] = (float)(0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (double)(float)n);
            mapd[n + 256] = 0.5 + 1.0 / 255.0 + 1.0 / 255.0 * (double)n;
        }
        mapd[254 + 256] = 1.5; mapf[254 + 256] = 1.5f;          // dither.cpp:90,102
        mapd[255 + 256] = 0.5 + 1.0 / 255.0 + 1.0 / 255.0 * 255.0; // out of bounds in the reference
        mapf[255 + 256] = (float)mapd[255 + 256];
        st.resize(n_channels);
        for (int n = 0; n < n_channels; n++) { st[n].randtab_ptr = n * spacing + 1; st[n].base = 0; }
    }

    void preloop(dither_state *s, int samples_per_loop) // dither.cpp:127-139
    {
        if (s->randtab_ptr + samples_per_loop >= size) {
            tab[0] = tab[s->randtab_ptr - 1];
            s->randtab_ptr = 1;
        }
        s->base = s->randtab_ptr;
        s->randtab_ptr += samples_per_loop;
    }
};

// quantisers: dither.cpp:215-274 / 349-409 (no dither), :141-212 / 276-347 (hp tpdf)
template <class T>
inline int32_t real2int_no_dither(T x, T rmin, T rmax, int32_t imin, int32_t imax, overflow_t *of)
{
    int32_t s;
    x += (T)0.5;
    if (x < 0) {
        if (x <= rmin) { s = imin; of->n_overflows++; if (x < -of->largest) of->largest = (double)-x; }
        else { s = (int32_t)x; s--; if (s < -of->intlargest) of->intlargest = -s; }
    } else {
        if (x > rmax) { s = imax; of->n_overflows++; if (x > of->largest) of->largest = (double)x; }
        else { s = (int32_t)x; if (s > of->intlargest) of->intlargest = s; }
    }
    return s;
}

template <class T>
inline int32_t real2int_hp_tpdf(T x, T rmin, T rmax, int32_t imin, int32_t imax, overflow_t *of,
                                T err[2], T dith)
{
    int32_t s;
    x += err[0] - err[1];
    err[1] = err[0];
    T d = x + dith;
    if (d < 0) {
        if (d <= rmin) { s = imin; of->n_overflows++; if (x < -of->largest) of->largest = (double)-d; }
        else { s = (int32_t)d; s--; if (s < -of->intlargest) of->intlargest = -s; }
    } else {
        if (d > rmax) { s = imax; of->n_overflows++; if (x > of->largest) of->largest = (double)d; }
        else { s = (int32_t)d; if (s > of->intlargest) of->intlargest = s; }
    }
    err[0] = x - (T)s;
    return s;
}

// real2raw.cpp:39-1221 (all four variants). `dith`==NULL -> *_no_dither.
template <class T>
void real2raw(uint8_t *raw, const T *real, const sample_format &sf, int spacing, int n_samples,
              overflow_t *of, Dither *dith, dither_state *ds, T err[2])
{
    const int stride = spacing * sf.bytes;
    if (sf.isfloat) { // REAL_OVERFLOW_UPDATE, real2raw.cpp:17-32
        const T rmin = (T)-of->max, rmax = (T)of->max;
        for (int n = 0; n < n_samples; n++, raw += stride) {
            T x = real[n];
            if (x < 0.0) {
                if (x < rmin) of->n_overflows++;
                if (x < -of->largest) of->largest = -x;
            } else {
                if (x > rmax) of->n_overflows++;
                if (x > of->largest) of->largest = x;
            }
            uint8_t b[8];
            if (sf.bytes == 4) { float f = (float)x; memcpy(b, &f, 4); }
            else { double d = (double)x; memcpy(b, &d, 8); }
            for (int i = 0; i < sf.bytes; i++) raw[i] = sf.swap ? b[sf.bytes - 1 - i] : b[i];
        }
        return;
    }
    const int bits = sf.bytes << 3;
    const int32_t imin = -(int32_t)(1u << (bits - 1)), imax = (int32_t)((1u << (bits - 1)) - 1);
    const T rmin = (T)imin, rmax = (T)imax;
    for (int n = 0; n < n_samples; n++, raw += stride) {
        int32_t s;
        if (dith != NULL) {
            int d = (int)dith->tab[ds->base + n] - (int)dith->tab[ds->base + n - 1];
            T dv = sizeof(T) == 4 ? (T)dith->mapf[d + 256] : (T)dith->mapd[d + 256];
            s = real2int_hp_tpdf<T>(real[n], rmin, rmax, imin, imax, of, err, dv);
        } else {
            s = real2int_no_dither<T>(real[n], rmin, rmax, imin, imax, of);
        }
        uint32_t u = (uint32_t)s;
        for (int i = 0; i < sf.bytes; i++) {
            uint8_t byte = (uint8_t)(u >> (8 * i));
            raw[sf.swap ? sf.bytes - 1 - i : i] = byte;
        }
    }
}

// ------------------------------------------------------------------ convolver (fftw_convolver.cpp)
template <class T>
struct Conv {
    int L, N;
    oracle_fft::RealFFT<T> fft;
    std::vector<T> scratch;

    explicit Conv(int length) : L(length), N(2 * length), fft(2 * length), scratch(2 * length) {}

    // :859-1158 / :1559-1855, MIXMODE_INPUT: HC -> ORD with scale and mix
    void mix_input(T *const *in, T *out, const double *scales, int n_bufs)
    {
        const int half = N >> 1;
        std::vector<T> tmp(N);
        for (int n = 0; n < half; n += 4)
            for (int j = 0; j < 4; j++) {
                T acc = in[0][n + j] * (T)scales[0];
                for (int i = 1; i < n_bufs; i++) acc += in[i][n + j] * (T)scales[i];
                tmp[(n << 1) + j] = acc;
            }
        {   // group 0 imaginary lanes: Nyquist, Im X1..X3 (:892-896)
            T acc = in[0][half] * (T)scales[0];
            for (int i = 1; i < n_bufs; i++) acc += in[i][half] * (T)scales[i];
            tmp[4] = acc;
            for (int j = 1; j < 4; j++) {
                acc = in[0][N - j] * (T)scales[0];
                for (int i = 1; i < n_bufs; i++) acc += in[i][N - j] * (T)scales[i];
                tmp[4 + j] = acc;
            }
        }
        for (int n = 4; n < half; n += 4)
            for (int j = 0; j < 4; j++) {
                T acc = in[0][N - n - j] * (T)scales[0];
                for (int i = 1; i < n_bufs; i++) acc += in[i][N - n - j] * (T)scales[i];
                tmp[(n << 1) + 4 + j] = acc;
            }
        memcpy(out, tmp.data(), sizeof(T) * N); // tmp: out may alias an input
    }

    // :1160-1421 / :1857-2117, MIXMODE_OUTPUT: ORD -> HC with scale and mix
    void mix_output(T *const *in, T *out, const double *scales, int n_bufs)
    {
        const int half = N >> 1;
        std::vector<T> tmp(N);
        for (int n = 0; n < half; n += 4)
            for (int j = 0; j < 4; j++) {
                T acc = in[0][(n << 1) + j] * (T)scales[0];
                for (int i = 1; i < n_bufs; i++) acc += in[i][(n << 1) + j] * (T)scales[i];
                tmp[n + j] = acc;
            }
        {
            T acc = in[0][4] * (T)scales[0];
            for (int i = 1; i < n_bufs; i++) acc += in[i][4] * (T)scales[i];
            tmp[half] = acc;
            for (int j = 1; j < 4; j++) {
                acc = in[0][4 + j] * (T)scales[0];
                for (int i = 1; i < n_bufs; i++) acc += in[i][4 + j] * (T)scales[i];
                tmp[N - j] = acc;
            }
        }
        for (int n = 4; n < half; n += 4)
            for (int j = 0; j < 4; j++) {
                T acc = in[0][(n << 1) + 4 + j] * (T)scales[0];
                for (int i = 1; i < n_bufs; i++) acc += in[i][(n << 1) + 4 + j] * (T)scales[i];
                tmp[N - n - j] = acc;
            }
        memcpy(out, tmp.data(), sizeof(T) * N);
    }

    // :1465-1493 / :2161-2189
    void convolve(const T *b, const T *c, T *d)
    {
        T d1s = b[0] * c[0], d2s = b[4] * c[4];
        for (int n = 0; n < N; n += 8)
            for (int j = 0; j < 4; j++) {
                T re = b[n + j] * c[n + j] - b[n + 4 + j] * c[n + 4 + j];
                T im = b[n + j] * c[n + 4 + j] + b[n + 4 + j] * c[n + j];
                d[n + j] = re;
                d[n + 4 + j] = im;
            }
        d[0] = d1s;
        d[4] = d2s;
    }

    // :1497-1525 / :2192-2220
    void convolve_add(const T *b, const T *c, T *d)
    {
        T d1s = d[0] + b[0] * c[0], d2s = d[4] + b[4] * c[4];
        for (int n = 0; n < N; n += 8)
            for (int j = 0; j < 4; j++) {
                d[n + j] += b[n + j] * c[n + j] - b[n + 4 + j] * c[n + 4 + j];
                d[n + 4 + j] += b[n + j] * c[n + 4 + j] + b[n + 4 + j] * c[n + j];
            }
        d[0] = d1s;
        d[4] = d2s;
    }

    // :1528-1556 / :2223-2251 -- on HC layout, sign by raw index parity
    void dirac(const T *in, T *out)
    {
        T fraction = (T)(1.0 / (T)N);
        for (int n = 0; n < N; n += 2) {
            out[n] = in[n] * +fraction;
            out[n + 1] = in[n + 1] * -fraction;
        }
    }

    // :475-537. Returns false on NaN/Inf.
    bool coeffs2cbuf(const T *coeffs, int n_coeffs, double scale, T *dest)
    {
        int len = n_coeffs > L ? L : n_coeffs;
        std::vector<T> r(N, (T)0);
        for (int n = 0; n < len; n++) {
            r[L + n] = coeffs[n] * (T)scale;
            if (!std::isfinite((double)r[L + n])) return false;
        }
        fft.r2hc(r.data(), r.data());
        double s = 1.0 / (double)N;
        T *in = r.data();
        mix_input(&in, dest, &s, 1);
        return true;
    }

    // :540-567
    void runtime_coeffs2cbuf(const T *src, T *dest)
    {
        memset(dest, 0, sizeof(T) * L);
        memmove(dest + L, src, sizeof(T) * L);
        fft.r2hc(dest, scratch.data());
        double s = 1.0 / (double)N;
        T *in = scratch.data();
        mix_input(&in, dest, &s, 1);
    }

    // :276-321, float-branch algorithm in both precisions
    void crossfade_inplace(T *input, T *xfade, T *buffer)
    {
        double one = 1.0;
        mix_output(&xfade, buffer, &one, 1);
        fft.hc2r(buffer, xfade);
        mix_output(&input, buffer, &one, 1);
        fft.hc2r(buffer, buffer);
        if (sizeof(T) == 4) {
            float f = (float)(1.0 / (float)(L - 1));
            for (int n = 0; n < L; n++) // operands promoted to double by the 1.0 literal (:301-303)
                buffer[n] = (T)((float)xfade[n] * (1.0 - f * (float)n) + (float)buffer[n] * f * (float)n);
        } else {
            double d = 1.0 / (double)(L - 1);
            for (int n = 0; n < L; n++)
                buffer[n] = (T)((double)xfade[n] * (1.0 - d * (double)n) + (double)buffer[n] * d * (double)n);
        }
        fft.r2hc(buffer, buffer);
        double s = 1.0 / (double)N;
        mix_input(&buffer, input, &s, 1);
    }

    // :378-403 (HC buffers; buffer is 1.5 N)
    void convolve_eval(const T *in, T *buffer, T *out)
    {
        fft.hc2r(in, buffer + L);
        fft.r2hc(buffer, out);
        memcpy(buffer, buffer + L, sizeof(T) * L);
    }
};

// ------------------------------------------------------------------ engine (brutefir.cpp)
template <class T>
struct Engine {
    int L, N, P, C;
    sample_format in_sf, out_sf;
    bool apply_dither;
    Conv<T> conv;
    Dither dith;
    std::vector<std::vector<T> > fdl;   // cbuf[n][P] (brutefir.cpp:775-781)
    std::vector<std::vector<T> > coeffs; // bfconf->coeffs[n].data[i]
    std::vector<int> coeff_blocks;
    std::vector<T> tprev;               // previous block per channel (input_timecbuf, :796-800)
    std::vector<int> procblocks;
    std::vector<overflow_t> overflow;
    std::vector<T> err;                 // dither_state_t.sf / .sd
    unsigned int blockcounter;
    bool initialized;

    Engine(int length, int blocks, int channels, int in_format, int out_format, int rate, bool dither_on)
        : L(length), N(2 * length), P(blocks), C(channels), apply_dither(dither_on), conv(length),
          dith(channels, rate, (int)sizeof(T), 0, length), blockcounter(0), initialized(false)
    {
        fill_format(&in_sf, in_format, true);
        fill_format(&out_sf, out_format, false);
        fdl.assign(C, std::vector<T>((size_t)P * N, (T)0));
        coeffs.assign(C, std::vector<T>());
        coeff_blocks.assign(C, 0);
        tprev.assign((size_t)C * L, (T)0);
        overflow.resize(C);
        err.assign((size_t)C * 2, (T)0);
        procblocks.assign(C, 0);
        for (int n = 0; n < C; n++) { // brutefir.cpp:669-684
            memset(&overflow[n], 0, sizeof(overflow_t));
            overflow[n].max = out_sf.isfloat ? 1.0 : (double)(1 << ((out_sf.bytes << 3) - 1)) - 1;
        }
    }

    void reset() // brutefir.cpp:347-367 (buffers are NOT cleared)
    {
        for (int n = 0; n < C; n++) { overflow[n].n_overflows = 0; overflow[n].largest = 0; overflow[n].intlargest = 0; }
        procblocks.assign(C, 0);
        blockcounter = 0;
    }

    // brutefir.cpp:180-228 + coeff.cpp:293-354
    int set_coeff(void *const *c, int n_coeffs, int length, int blocks, double scale)
    {
        initialized = false;
        for (int n = 0; n < C; n++) { coeffs[n].clear(); coeff_blocks[n] = 0; }
        if (n_coeffs > C) n_coeffs = C;
        std::vector<T> zero(L, (T)0);
        for (int n = 0; n < n_coeffs; n++) {
            coeffs[n].assign((size_t)blocks * N, (T)0);
            const T *src = (const T *)c[n];
            for (int i = 0; i < blocks; i++) {
                bool ok;
                if (i * L > length) ok = conv.coeffs2cbuf(zero.data(), L, scale, &coeffs[n][(size_t)i * N]);
                else if ((i + 1) * L > length) ok = conv.coeffs2cbuf(src + (size_t)i * L, length - i * L, scale, &coeffs[n][(size_t)i * N]);
                else ok = conv.coeffs2cbuf(src + (size_t)i * L, L, scale, &coeffs[n][(size_t)i * N]);
                if (!ok) { for (int k = 0; k < C; k++) { coeffs[k].clear(); coeff_blocks[k] = 0; } return -2; }
            }
            coeff_blocks[n] = blocks;
        }
        initialized = true;
        return 0;
    }

    // brutefir.cpp:245-343
    int run(const void *inbuf, void *outbuf)
    {
        std::vector<T> tcur(N), freq(N), acc(N), hc(N), y(N);
        for (int n = 0; n < C; n++) {
            if (coeff_blocks[n] == 0) return -3; // the reference would dereference NULL here
            // raw2cbuf (:255-260 -> fftw_convolver.cpp:157-185): [prev | cur]
            memcpy(tcur.data(), &tprev[(size_t)n * L], sizeof(T) * L);
            raw2real<T>(tcur.data() + L, (const uint8_t *)inbuf + n * in_sf.bytes, in_sf, C, L);
            memcpy(&tprev[(size_t)n * L], tcur.data() + L, sizeof(T) * L);
            conv.fft.r2hc(tcur.data(), freq.data());                       // :263
            if (procblocks[n] < P) procblocks[n]++;                        // :265-268
            int curblock = (int)(blockcounter % (unsigned int)P);          // :270
            T *fin = freq.data();
            conv.mix_input(&fin, &fdl[n][(size_t)curblock * N], &in_sf.scale, 1); // :273-277
            if (P == 1) {                                                  // :279-284
                conv.convolve(&fdl[n][0], &coeffs[n][0], acc.data());
            } else {
                conv.convolve(&fdl[n][(size_t)curblock * N], &coeffs[n][0], acc.data()); // :288-290
                for (int i = 1; i < coeff_blocks[n] && i < procblocks[n]; i++) {           // :292-299
                    int convblock = (int)((blockcounter - i) % (unsigned int)P);
                    conv.convolve_add(&fdl[n][(size_t)convblock * N], &coeffs[n][(size_t)i * N], acc.data());
                }
            }
            T *ain = acc.data();
            conv.mix_output(&ain, hc.data(), &out_sf.scale, 1);            // :303-307
            conv.fft.hc2r(hc.data(), y.data());                            // :311
            if (!std::isfinite((double)y[0])) return -1;                   // :316-321
            overflow_t of = overflow[n];                                   // :324-333 -> fftw_convolver.cpp:406-466
            bool dither_now = apply_dither && !out_sf.isfloat;
            if (dither_now) dith.preloop(&dith.st[n], L);
            real2raw<T>((uint8_t *)outbuf + n * out_sf.bytes, y.data(), out_sf, C, L, &of,
                        dither_now ? &dith : NULL, &dith.st[n], &err[(size_t)n * 2]);
            overflow[n] = of;
        }
        blockcounter++; // :337-340
        return 0;
    }
};

// ---- small one-shot convolver on plain half-complex buffers (fftw_convolver.cpp:698-777, 820-856)
template <class T>
struct TdConv {
    int blocklen;
    oracle_fft::RealFFT<T> fft;
    std::vector<T> coeffs; // spectrum of [0_blocklen | h | 0], scaled by 1/(2 blocklen), HC layout
    TdConv(const T *h, int n, int bl) : blocklen(bl), fft(2 * bl), coeffs(2 * (size_t)bl, (T)0)
    {
        memcpy(&coeffs[bl], h, sizeof(T) * n);             // :731-736
        fft.r2hc(coeffs.data(), coeffs.data());            // :740 / :750
        const T s = (T)(1.0 / (T)(bl << 1));               // :741 / :751
        for (int i = 0; i < 2 * bl; i++) coeffs[i] *= s;
    }
    void convolve(T *b) // :763-777 with convolve_inplace_ordered :820-856
    {
        const int size = 2 * blocklen, size2 = blocklen;
        const T *c = coeffs.data();
        fft.r2hc(b, b);
        b[0] *= c[0];
        for (int n = 1; n < size2; n++) {
            const T a = b[n];
            b[n] = a * c[n] - b[size - n] * c[size - n];
            b[size - n] = a * c[size - n] + b[size - n] * c[n];
        }
        b[size2] *= c[size2];
        fft.hc2r(b, b);
    }
};
struct td_handle { int realsize; TdConv<float> *f; TdConv<double> *d; };

static int td_block_length(int n_coeffs) // :698-706 with log2_roof (log2.h:34-51); n == 1 is refused (shift by -1 there)
{
    if (n_coeffs < 2) return -1;
    int lg = 31;
    while (((unsigned)n_coeffs & (1u << lg)) == 0 && lg > 0) lg--;
    if (((unsigned)n_coeffs & ~(1u << lg)) != 0) lg++;
    return 1 << lg;
}
struct conv_handle {
    int realsize;
    Conv<float> *cf;
    Conv<double> *cd;
    Dither *dith;
    std::vector<float> errf;
    std::vector<double> errd;
};

struct engine_handle {
    int realsize;
    Engine<float> *ef;
    Engine<double> *ed;
};

// equalizer.cpp:183-204
template <class T> inline T cosine_int(T mag1, T mag2, T freq1, T freq2, T curfreq)
{
    return (T)((mag1 - mag2) * 0.5 * cos(M_PI * (curfreq - freq1) / (freq2 - freq1)) + (mag1 + mag2) * 0.5);
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- convolver handles (same shapes as oracle/ref_shim/ref_capi.cpp)
void *orc_conv_new(int length, int realsize, int n_channels, int sample_rate)
{
    if ((realsize != 4 && realsize != 8) || length < 1 || (length & (length - 1)) != 0) return NULL;
    conv_handle *h = new conv_handle;
    h->realsize = realsize;
    h->cf = realsize == 4 ? new Conv<float>(length) : NULL;
    h->cd = realsize == 8 ? new Conv<double>(length) : NULL;
    if (n_channels < 1) n_channels = 1;
    h->dith = new Dither(n_channels, sample_rate, realsize, 0, length);
    h->errf.assign((size_t)n_channels * 2, 0.0f);
    h->errd.assign((size_t)n_channels * 2, 0.0);
    return h;
}

void orc_conv_delete(void *p)
{
    conv_handle *h = (conv_handle *)p;
    delete h->cf; delete h->cd; delete h->dith; delete h;
}

int orc_conv_cbufsize(void *p)
{
    conv_handle *h = (conv_handle *)p;
    return h->realsize * (h->realsize == 4 ? h->cf->N : h->cd->N); // fftw_convolver.cpp:469-472
}

int orc_conv_raw2cbuf(void *p, void *rawbuf, void *cbuf, void *next_cbuf, int format, int index, int spacing)
{
    conv_handle *h = (conv_handle *)p; // fftw_convolver.cpp:157-185
    sample_format sf;
    if (fill_format(&sf, format, true) != 0) return -1;
    const uint8_t *raw = (const uint8_t *)rawbuf + index * sf.bytes;
    if (h->realsize == 4) {
        int L = h->cf->L;
        raw2real<float>((float *)next_cbuf, raw, sf, spacing, L);
        memcpy((float *)cbuf + L, next_cbuf, sizeof(float) * L);
    } else {
        int L = h->cd->L;
        raw2real<double>((double *)next_cbuf, raw, sf, spacing, L);
        memcpy((double *)cbuf + L, next_cbuf, sizeof(double) * L);
    }
    return 0;
}

void orc_conv_time2freq(void *p, void *in, void *out)
{
    conv_handle *h = (conv_handle *)p; // fftw_convolver.cpp:188-212
    if (h->realsize == 4) h->cf->fft.r2hc((float *)in, (float *)out); else h->cd->fft.r2hc((double *)in, (double *)out);
}

void orc_conv_freq2time(void *p, void *in, void *out)
{
    conv_handle *h = (conv_handle *)p; // fftw_convolver.cpp:351-375
    if (h->realsize == 4) h->cf->fft.hc2r((float *)in, (float *)out); else h->cd->fft.hc2r((double *)in, (double *)out);
}

void orc_conv_mixnscale(void *p, void **in, void *out, double *scales, int n_bufs, int mixmode)
{
    conv_handle *h = (conv_handle *)p; // fftw_convolver.cpp:215-229; mode 2 is unimplemented there too
    if (mixmode == MIX_INPUT) {
        if (h->realsize == 4) h->cf->mix_input((float *const *)in, (float *)out, scales, n_bufs);
        else h->cd->mix_input((double *const *)in, (double *)out, scales, n_bufs);
    } else if (mixmode == MIX_OUTPUT) {
        if (h->realsize == 4) h->cf->mix_output((float *const *)in, (float *)out, scales, n_bufs);
        else h->cd->mix_output((double *const *)in, (double *)out, scales, n_bufs);
    }
}

void orc_conv_convolve(void *p, void *in, void *coeffs, void *out)
{
    conv_handle *h = (conv_handle *)p;
    if (h->realsize == 4) h->cf->convolve((float *)in, (float *)coeffs, (float *)out);
    else h->cd->convolve((double *)in, (double *)coeffs, (double *)out);
}

void orc_conv_convolve_inplace(void *p, void *cbuf, void *coeffs)
{
    // fftw_convolver.cpp:1430-1462: same arithmetic as convolve with out == in (each group of 8 is
    // read before it is written)
    orc_conv_convolve(p, cbuf, coeffs, cbuf);
}

void orc_conv_convolve_add(void *p, void *in, void *coeffs, void *out)
{
    conv_handle *h = (conv_handle *)p;
    if (h->realsize == 4) h->cf->convolve_add((float *)in, (float *)coeffs, (float *)out);
    else h->cd->convolve_add((double *)in, (double *)coeffs, (double *)out);
}

void orc_conv_crossfade_inplace(void *p, void *in, void *xfade, void *buffer)
{
    conv_handle *h = (conv_handle *)p;
    if (h->realsize == 4) h->cf->crossfade_inplace((float *)in, (float *)xfade, (float *)buffer);
    else h->cd->crossfade_inplace((double *)in, (double *)xfade, (double *)buffer);
}

void orc_conv_dirac_convolve(void *p, void *in, void *out)
{
    conv_handle *h = (conv_handle *)p;
    if (h->realsize == 4) h->cf->dirac((float *)in, (float *)out); else h->cd->dirac((double *)in, (double *)out);
}

void orc_conv_dirac_convolve_inplace(void *p, void *cbuf) { orc_conv_dirac_convolve(p, cbuf, cbuf); }

void orc_conv_convolve_eval(void *p, void *in, void *buffer, void *out)
{
    conv_handle *h = (conv_handle *)p;
    if (h->realsize == 4) h->cf->convolve_eval((float *)in, (float *)buffer, (float *)out);
    else h->cd->convolve_eval((double *)in, (double *)buffer, (double *)out);
}

int orc_conv_td_block_length(void *, int n_coeffs) { return td_block_length(n_coeffs); }
void *orc_conv_td_new(void *p, void *coeffs, int n_coeffs)
{
    conv_handle *h = (conv_handle *)p;
    const int bl = td_block_length(n_coeffs);
    if (bl < 0) return NULL;
    td_handle *t = new td_handle();
    t->realsize = h->realsize;
    if (h->realsize == 4) t->f = new TdConv<float>((const float *)coeffs, n_coeffs, bl);
    else t->d = new TdConv<double>((const double *)coeffs, n_coeffs, bl);
    return t;
}
void orc_conv_td_coeffs(void *tdc, void *dst, int nbytes)
{
    td_handle *t = (td_handle *)tdc;
    memcpy(dst, t->realsize == 4 ? (const void *)t->f->coeffs.data() : (const void *)t->d->coeffs.data(), nbytes);
}
void orc_conv_td_convolve(void *, void *tdc, void *overlap_block)
{
    td_handle *t = (td_handle *)tdc;
    if (t->realsize == 4) t->f->convolve((float *)overlap_block); else t->d->convolve((double *)overlap_block);
}
void orc_conv_td_free(void *tdc)
{
    td_handle *t = (td_handle *)tdc;
    if (t == NULL) return;
    delete t->f; delete t->d; delete t;
}

int orc_conv_cbuf2raw(void *p, void *cbuf, void *outbuf, int format, int index, int spacing,
                      int apply_dither, int dither_channel, overflow_t *overflow)
{
    conv_handle *h = (conv_handle *)p; // fftw_convolver.cpp:406-466
    sample_format sf;
    if (fill_format(&sf, format, false) != 0) return -1;
    if (dither_channel < 0 || dither_channel >= (int)h->dith->st.size()) return -1;
    bool dither_now = apply_dither && !sf.isfloat;
    dither_state *ds = &h->dith->st[dither_channel];
    uint8_t *raw = (uint8_t *)outbuf + index * sf.bytes;
    if (h->realsize == 4) {
        if (dither_now) h->dith->preloop(ds, h->cf->L);
        real2raw<float>(raw, (const float *)cbuf, sf, spacing, h->cf->L, overflow, dither_now ? h->dith : NULL, ds, &h->errf[(size_t)dither_channel * 2]);
    } else {
        if (dither_now) h->dith->preloop(ds, h->cd->L);
        real2raw<double>(raw, (const double *)cbuf, sf, spacing, h->cd->L, overflow, dither_now ? h->dith : NULL, ds, &h->errd[(size_t)dither_channel * 2]);
    }
    return 0;
}

int orc_conv_coeffs2cbuf(void *p, void *coeffs, int n_coeffs, double scale, void *dest)
{
    conv_handle *h = (conv_handle *)p;
    bool ok = h->realsize == 4 ? h->cf->coeffs2cbuf((const float *)coeffs, n_coeffs, scale, (float *)dest)
                               : h->cd->coeffs2cbuf((const double *)coeffs, n_coeffs, scale, (double *)dest);
    return ok ? 0 : -1;
}

void orc_conv_runtime_coeffs2cbuf(void *p, void *src, void *dest)
{
    conv_handle *h = (conv_handle *)p;
    if (h->realsize == 4) h->cf->runtime_coeffs2cbuf((const float *)src, (float *)dest);
    else h->cd->runtime_coeffs2cbuf((const double *)src, (double *)dest);
}

int orc_conv_dither_table_size(void *p) { return ((conv_handle *)p)->dith->size; }
const int8_t *orc_conv_dither_table(void *p) { return ((conv_handle *)p)->dith->tab.data(); }
int orc_conv_dither_ptr(void *p, int ch) { return ((conv_handle *)p)->dith->st[ch].randtab_ptr; }
void orc_conv_dither_map(void *p, void *out)
{
    conv_handle *h = (conv_handle *)p; // 511 entries [-256..254] like the reference table
    if (h->realsize == 4) memcpy(out, h->dith->mapf.data(), 511 * sizeof(float));
    else memcpy(out, h->dith->mapd.data(), 511 * sizeof(double));
}

void orc_raw2real(int realsize, void *realbuf, void *rawbuf, int bytes, int shift, int isfloat, int spacing, int swap, int n)
{
    (void)shift; // always 0 for the formats of global.h:24-37 (bytes == sbytes)
    sample_format sf; sf.bytes = bytes; sf.isfloat = isfloat != 0; sf.swap = swap != 0; sf.scale = 1.0;
    if (realsize == 4) raw2real<float>((float *)realbuf, (const uint8_t *)rawbuf, sf, spacing, n);
    else raw2real<double>((double *)realbuf, (const uint8_t *)rawbuf, sf, spacing, n);
}

// ---------------------------------------------------------------- engine
void *orc_bfir_new(int filter_length, int filter_blocks, int realsize, int channels, int in_format,
                   int out_format, int sampling_rate, int apply_dither)
{
    sample_format sf;
    if (channels < 1 || filter_blocks < 1) return NULL;
    if ((realsize != 4 && realsize != 8) || filter_length < 1 || (filter_length & (filter_length - 1)) != 0) return NULL;
    if (fill_format(&sf, in_format, true) != 0 || fill_format(&sf, out_format, false) != 0) return NULL;
    engine_handle *h = new engine_handle;
    h->realsize = realsize;
    h->ef = realsize == 4 ? new Engine<float>(filter_length, filter_blocks, channels, in_format, out_format, sampling_rate, apply_dither != 0) : NULL;
    h->ed = realsize == 8 ? new Engine<double>(filter_length, filter_blocks, channels, in_format, out_format, sampling_rate, apply_dither != 0) : NULL;
    return h;
}

void orc_bfir_delete(void *p) { engine_handle *h = (engine_handle *)p; delete h->ef; delete h->ed; delete h; }
int orc_bfir_is_initialized(void *p) { engine_handle *h = (engine_handle *)p; return (h->realsize == 4 ? h->ef->initialized : h->ed->initialized) ? 1 : 0; }

int orc_bfir_set_coeff(void *p, void **coeffs, int n_coeffs, int length, int coeff_blocks, double scale)
{
    engine_handle *h = (engine_handle *)p;
    return h->realsize == 4 ? h->ef->set_coeff(coeffs, n_coeffs, length, coeff_blocks, scale)
                            : h->ed->set_coeff(coeffs, n_coeffs, length, coeff_blocks, scale);
}

int orc_bfir_run(void *p, void *inbuf, void *outbuf)
{
    engine_handle *h = (engine_handle *)p;
    return h->realsize == 4 ? h->ef->run(inbuf, outbuf) : h->ed->run(inbuf, outbuf);
}

void orc_bfir_reset(void *p) { engine_handle *h = (engine_handle *)p; if (h->realsize == 4) h->ef->reset(); else h->ed->reset(); }

void orc_bfir_get_overflow(void *p, int ch, overflow_t *out)
{
    engine_handle *h = (engine_handle *)p;
    *out = h->realsize == 4 ? h->ef->overflow[ch] : h->ed->overflow[ch];
}

int orc_bfir_dither_ptr(void *p, int ch)
{
    engine_handle *h = (engine_handle *)p;
    return h->realsize == 4 ? h->ef->dith.st[ch].randtab_ptr : h->ed->dith.st[ch].randtab_ptr;
}

unsigned int orc_bfir_blockcounter(void *p)
{
    engine_handle *h = (engine_handle *)p;
    return h->realsize == 4 ? h->ef->blockcounter : h->ed->blockcounter;
}

// coeff.cpp:293-354
int orc_preprocess_coeff(void *p, void *coeffs, int filter_length, int coeff_blocks, int coeff_length,
                         int realsize, double scale, void *dest)
{
    conv_handle *h = (conv_handle *)p;
    const int L = filter_length, N = 2 * L;
    int rc = 0;
    for (int n = 0; n < coeff_blocks; n++) {
        int len;
        const uint8_t *src = (const uint8_t *)coeffs + (size_t)n * L * realsize;
        std::vector<uint8_t> zero((size_t)L * realsize, 0);
        if (n * L > coeff_length) { src = zero.data(); len = L; }
        else if ((n + 1) * L > coeff_length) len = coeff_length - n * L;
        else len = L;
        if (orc_conv_coeffs2cbuf(p, (void *)src, len, scale, (uint8_t *)dest + (size_t)n * N * realsize) != 0) rc = -1;
    }
    (void)h;
    return rc;
}

// equalizer.cpp:30-67 (ctor), :87-140 (generate), :212-299 (render_f), :307-394 (render_d)
int orc_equalizer_render(int block_length, int n_blocks, int realsize, int n_channels, int sampling_rate,
                         int n_bands, double *freq, double *mag, double *phase, void *out, int out_len)
{
    static const double iso_bands[31] = { 20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500,
        630, 800, 1000, 1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500, 16000, 20000 };
    const int taps = block_length * n_blocks, bc = 33;
    (void)n_channels;
    if (taps < 2 || (taps & (taps - 1)) != 0 || n_bands > 31) return -1;
    double efreq[33], emag[33], ephase[33];
    memset(emag, 0, sizeof(emag)); memset(ephase, 0, sizeof(ephase));
    efreq[0] = 0.0; efreq[bc - 1] = (double)sampling_rate / 2.0;
    for (int n = 0; n < 31; n++) efreq[n + 1] = iso_bands[n];
    for (int n = 0, i = 0; n < n_bands; n++) {
        while (freq[n] > efreq[i]) i++;
        emag[i] = mag[n]; ephase[i] = phase[n]; i++;
    }
    emag[0] = emag[1]; emag[bc - 1] = emag[bc - 2];
    for (int n = 0; n < bc; n++) {
        efreq[n] /= (double)sampling_rate;
        emag[n] = pow(10, emag[n] / 20);
        ephase[n] /= (180 * M_PI); // sic, equalizer.cpp:120
    }
    int frames = (taps >> 1) < out_len ? (taps >> 1) : out_len;
    if (realsize == 4) {
        std::vector<float> r(taps);
        float fm[33], ff[33], fp[33];
        for (int n = 0; n < bc; n++) { fm[n] = (float)emag[n]; ff[n] = (float)efreq[n]; fp[n] = (float)ephase[n]; }
        float scale = (float)(1.0 / (float)taps), divtaps = (float)(1.0 / (float)taps);
        float tapspi = (float)(-(float)taps * M_PI);
        r[0] = fm[0] * scale;
        for (int n = 1, i = 0; n < taps >> 1; n++) {
            float curfreq = (float)n * divtaps;
            while (curfreq > ff[i + 1]) i++;
            float m = cosine_int<float>(fm[i], fm[i + 1], ff[i], ff[i + 1], curfreq) * scale;
            float rad = tapspi * curfreq + cosine_int<float>(fp[i], fp[i + 1], ff[i], ff[i + 1], curfreq);
            r[n] = (float)(cos(rad) * m);
            r[taps - n] = (float)(sin(rad) * m);
        }
        r[taps >> 1] = fm[bc - 1] * scale;
        oracle_fft::RealFFT<float> fft(taps);
        fft.hc2r(r.data(), r.data());
        memcpy(out, &r[taps >> 1], sizeof(float) * frames);
    } else {
        std::vector<double> r(taps);
        double scale = 1.0 / (double)taps, divtaps = 1.0 / (double)taps, tapspi = -(double)taps * M_PI;
        r[0] = emag[0] * scale;
        for (int n = 1, i = 0; n < taps >> 1; n++) {
            double curfreq = (double)n * divtaps;
            while (curfreq > efreq[i + 1]) i++;
            double m = cosine_int<double>(emag[i], emag[i + 1], efreq[i], efreq[i + 1], curfreq) * scale;
            double rad = tapspi * curfreq + cosine_int<double>(ephase[i], ephase[i + 1], efreq[i], efreq[i + 1], curfreq);
            r[n] = cos(rad) * m;
            r[taps - n] = sin(rad) * m;
        }
        r[taps >> 1] = emag[bc - 1] * scale;
        oracle_fft::RealFFT<double> fft(taps);
        fft.hc2r(r.data(), r.data());
        memcpy(out, &r[taps >> 1], sizeof(double) * frames);
    }
    return frames;
}

const char *orc_fft_provider(void) { return "oracle/fft_r2r (own Stockham radix-4, native precision; NOT FFTW)"; }

}
