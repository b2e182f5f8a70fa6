// Raw sample codecs: the per-sample arithmetic of the reference's raw2real / real2raw / dither
// (brutefir/raw2real.cpp:17-424, brutefir/real2raw.cpp:39-1221, brutefir/dither.cpp:141-409), written
// as __host__ __device__ functions that the FFT kernels call from their load / store phases.
// Sample format codes are the reference's (brutefir/global.h:24-37).
#pragma once
#include <stdint.h>
#include <string.h>
#include <cuda_runtime.h>

#ifndef BFIR_HD
#define BFIR_HD __host__ __device__ __forceinline__
#endif

namespace bfir {

enum {
    FMT_S8 = 1, FMT_S16_LE = 2, FMT_S16_BE = 3, FMT_S24_LE = 4, FMT_S24_BE = 5, FMT_S32_LE = 6, FMT_S32_BE = 7,
    FMT_FLOAT_LE = 8, FMT_FLOAT_BE = 9, FMT_FLOAT64_LE = 10, FMT_FLOAT64_BE = 11
};

BFIR_HD int fmt_bytes(int fmt)
{
    switch (fmt) {
    case FMT_S8: return 1;
    case FMT_S16_LE: case FMT_S16_BE: return 2;
    case FMT_S24_LE: case FMT_S24_BE: return 3;
    case FMT_S32_LE: case FMT_S32_BE: case FMT_FLOAT_LE: case FMT_FLOAT_BE: return 4;
    case FMT_FLOAT64_LE: case FMT_FLOAT64_BE: return 8;
    default: return 0;
    }
}
BFIR_HD bool fmt_isfloat(int fmt) { return fmt >= FMT_FLOAT_LE && fmt <= FMT_FLOAT64_BE; }
BFIR_HD bool fmt_valid(int fmt) { return fmt >= FMT_S8 && fmt <= FMT_FLOAT64_BE; }

BFIR_HD uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
BFIR_HD uint64_t bswap64(uint64_t v) { return ((uint64_t)bswap32((uint32_t)v) << 32) | bswap32((uint32_t)(v >> 32)); }

BFIR_HD float u32_as_float(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
BFIR_HD uint32_t float_as_u32(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
BFIR_HD double u64_as_double(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
BFIR_HD uint64_t double_as_u64(double d)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

// One raw sample -> real. p must be aligned to the sample size for the 2/4/8-byte formats (it always
// is: byte_offset = channel * bytes and spacing counts whole samples, brutefir.cpp:557-558).
template <class T> BFIR_HD T load_raw(const uint8_t *p, int fmt)
{
    switch (fmt) {
    case FMT_S8: return (T)(*(const int8_t *)p);                                          // raw2real.cpp:86-91
    case FMT_S16_LE: return (T)(*(const int16_t *)p);                                     // :101-107
    case FMT_S16_BE: { uint16_t u = *(const uint16_t *)p; return (T)(int16_t)((u >> 8) | (u << 8)); } // :94-100
    case FMT_S24_LE: // bytes into the top of an int32, arithmetic >> 8 (:165-174)
        return (T)((int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24)) >> 8);
    case FMT_S24_BE: // :155-164
        return (T)((int32_t)(((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24)) >> 8);
    case FMT_S32_LE: return (T)(*(const int32_t *)p);                                     // :188-194
    case FMT_S32_BE: return (T)(int32_t)bswap32(*(const uint32_t *)p);                    // :181-187
    case FMT_FLOAT_LE: return (T)(*(const float *)p);                                     // :50-56
    case FMT_FLOAT_BE: return (T)u32_as_float(bswap32(*(const uint32_t *)p));             // :43-49
    case FMT_FLOAT64_LE: return (T)(*(const double *)p);                                  // :68-74
    case FMT_FLOAT64_BE: return (T)u64_as_double(bswap64(*(const uint64_t *)p));          // :60-67
    default: return (T)0;
    }
}

// per-thread overflow statistics; merged into the channel's bfoverflow_t image by the caller
struct OverflowAcc {
    unsigned int n_overflows;
    int32_t intlargest;
    double largest;
};

struct OverflowStats { // device image of bfoverflow_t (global.h:96-102) without `max`
    unsigned int n_overflows;
    int32_t intlargest;
    unsigned long long largest_bits; // non-negative double, compared as an integer by atomicMax
};

// float / double output with the REAL_OVERFLOW_UPDATE statistics (real2raw.cpp:17-32)
template <class T> BFIR_HD void store_raw_float(uint8_t *p, int fmt, T x, T rmax, OverflowAcc &acc)
{
    if (x < (T)0) {
        if (x < -rmax) acc.n_overflows++;
        if ((double)x < -acc.largest) acc.largest = -(double)x;
    } else {
        if (x > rmax) acc.n_overflows++;
        if ((double)x > acc.largest) acc.largest = (double)x;
    }
    switch (fmt) {
    case FMT_FLOAT_LE: *(float *)p = (float)x; break;
    case FMT_FLOAT_BE: *(uint32_t *)p = bswap32(float_as_u32((float)x)); break;
    case FMT_FLOAT64_LE: *(double *)p = (double)x; break;
    default: *(uint64_t *)p = bswap64(double_as_u64((double)x)); break;
    }
}

template <class T> BFIR_HD int32_t real_to_int_rz(T x)
{
#ifdef __CUDA_ARCH__
    return sizeof(T) == 4 ? __float2int_rz((float)x) : __double2int_rz((double)x);
#else
    return (int32_t)x;
#endif
}

// dither.cpp:215-274 / :349-409. `x` is the sample AFTER the caller added 0.5 (no dither) or the
// dither value; `stat` is the value the statistics compare (identical to x without dither).
template <class T> BFIR_HD int32_t quantise(T x, T stat, T rmin, T rmax, int32_t imin, int32_t imax, OverflowAcc &acc)
{
    int32_t s;
    if (x < (T)0) {
        if (x <= rmin) {
            s = imin;
            acc.n_overflows++;
            if ((double)stat < -acc.largest) acc.largest = -(double)x;
        } else {
            s = real_to_int_rz<T>(x);
            s--;
            if (s < -acc.intlargest) acc.intlargest = -s;
        }
    } else {
        if (x > rmax) {
            s = imax;
            acc.n_overflows++;
            if ((double)stat > acc.largest) acc.largest = (double)x;
        } else {
            s = real_to_int_rz<T>(x);
            if (s > acc.intlargest) acc.intlargest = s;
        }
    }
    return s;
}

// Branch-free forms of the two statistics updates above for the unrolled store loops of the transform kernels
// (there every branch costs its resolve latency sixteen times per thread, and `largest` as a double costs a
// conversion and an FP64 compare per sample in single-precision engines): the per-thread maximum is kept in T --
// exact, |x| converts to double without rounding and max does not round -- and merged into OverflowAcc once.
// Same counters as store_raw_float / quantise for every input, NaN included (all comparisons false).
template <class T> BFIR_HD void float_stats_nb(T x, T rmax, T &largest, unsigned int &n_overflows)
{
    const T ax = x < (T)0 ? -x : x;
    n_overflows += ax > rmax ? 1u : 0u;
    largest = ax > largest ? ax : largest;
}

template <class T> BFIR_HD int32_t quantise_nb(T x, T rmin, T rmax, int32_t imin, int32_t imax, T &largest, unsigned int &n_overflows, int32_t &intlargest)
{
    const bool neg = x < (T)0;
    const bool ovn = neg && x <= rmin, ovp = !neg && x > rmax, ov = ovn || ovp;
    int32_t s = real_to_int_rz<T>(x) - (neg ? 1 : 0);
    s = ovn ? imin : (ovp ? imax : s);
    const T ax = neg ? -x : x;
    n_overflows += ov ? 1u : 0u;
    largest = (ov && ax > largest) ? ax : largest;
    const int32_t as = neg ? (int32_t)(0u - (uint32_t)s) : s;
    intlargest = (!ov && as > intlargest) ? as : intlargest;
    return s;
}

BFIR_HD void store_raw_int(uint8_t *p, int fmt, int32_t s)
{
    const uint32_t u = (uint32_t)s;
    switch (fmt) {
    case FMT_S8: *(int8_t *)p = (int8_t)s; break;
    case FMT_S16_LE: *(int16_t *)p = (int16_t)s; break;
    case FMT_S16_BE: *(uint16_t *)p = (uint16_t)(((u & 0xff) << 8) | ((u >> 8) & 0xff)); break;
    case FMT_S24_LE: p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); break;
    case FMT_S24_BE: p[0] = (uint8_t)(u >> 16); p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)u; break;
    case FMT_S32_LE: *(int32_t *)p = s; break;
    default: *(uint32_t *)p = bswap32(u); break;
    }
}

BFIR_HD void int_limits(int fmt, int32_t &imin, int32_t &imax)
{
    const int bits = fmt_bytes(fmt) << 3;
    imin = (int32_t)(0u - (1u << (bits - 1)));
    imax = (int32_t)((1u << (bits - 1)) - 1u);
}

// x + 0.5 without contraction (the no-dither requantiser's rounding offset)
template <class T> BFIR_HD T half_up(T x)
{
#ifdef __CUDA_ARCH__
    return sizeof(T) == 4 ? (T)__fadd_rn((float)x, 0.5f) : (T)__dadd_rn((double)x, 0.5);
#else
    return x + (T)0.5;
#endif
}

// float / double sample in the raw format FMT (compile-time), no statistics
template <class T, int FMT> BFIR_HD void store_raw_real(uint8_t *p, T x)
{
    if (FMT == FMT_FLOAT_LE) *(float *)p = (float)x;
    else if (FMT == FMT_FLOAT_BE) *(uint32_t *)p = bswap32(float_as_u32((float)x));
    else if (FMT == FMT_FLOAT64_LE) *(double *)p = (double)x;
    else *(uint64_t *)p = bswap64(double_as_u64((double)x));
}

// integer output without dither: real2raw*_no_dither -> dither*_real2int_no_dither
template <class T> BFIR_HD void store_raw_quantised(uint8_t *p, int fmt, T x, T rmin, T rmax, int32_t imin, int32_t imax, OverflowAcc &acc)
{
#ifdef __CUDA_ARCH__
    const T xh = sizeof(T) == 4 ? (T)__fadd_rn((float)x, 0.5f) : (T)__dadd_rn((double)x, 0.5);
#else
    const T xh = x + (T)0.5;
#endif
    store_raw_int(p, fmt, quantise<T>(xh, xh, rmin, rmax, imin, imax, acc));
}

} // namespace bfir
