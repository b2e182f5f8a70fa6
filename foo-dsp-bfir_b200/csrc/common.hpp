// Host-side helpers shared by the engine, the convolver and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>
#include <stdint.h>
#include "../../include/bfir_b200.h"
#include "fft_dispatch.hpp"
#include "mac_kernels.cuh"
#include "elementwise_kernels.cuh"

namespace bfir {

extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_launches;
extern void (*g_print_cb)(const char *);

void set_error(const char *fmt, ...);
void pinfo(const char *fmt, ...); // reference brutefir/pinfo.c:26-39

extern thread_local unsigned long long t_launches;   // this thread's share (graph capture counts its own launches with it)
inline void count_launch(unsigned long long n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); t_launches += n; }

#define BFIR_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t err__ = (call);                                                              \
        if (err__ != cudaSuccess) {                                                              \
            bfir::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return BFIR_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

inline int ilog2_exact(int x)
{
    if (x < 1 || (x & (x - 1)) != 0) return -1; // log2_get, reference brutefir/log2.h:18-31
    int lg = 0;
    while ((1 << lg) < x) lg++;
    return lg;
}

// sample_format_t as filled by brutefir::setup_sample_format, brutefir.cpp:436-539
struct SampleFormat {
    int format, bytes;
    bool isfloat;
    double scale;
};
inline bool fill_sample_format(SampleFormat *sf, int format, bool normalized)
{
    if (!fmt_valid(format)) return false;
    sf->format = format;
    sf->bytes = fmt_bytes(format);
    sf->isfloat = fmt_isfloat(format);
    if (sf->isfloat) sf->scale = 1.0;
    else {
        const double full = (double)(1 << ((sf->bytes << 3) - 1)); // get_full_scale, brutefir.cpp:397-401
        sf->scale = normalized ? 1.0 / full : full;
    }
    return true;
}

// exp(-2 pi i j / n), j = 0..n-1, rounded once from long double; device table of cpx<T>
int make_twiddles(int realsize, int n, void **d_out);

// The reference's dither object (brutefir/dither.cpp:21-110): Tausworthe byte table + TPDF map +
// per-channel start offsets, uploaded to the device.
struct DitherTables {
    std::vector<int8_t> h_tab;
    int size = 0, spacing = 0, realsize = 0, n_channels = 0;
    int8_t *d_tab = nullptr;
    void *d_map = nullptr;          // 512 entries of T: randmap[-256..255]
    DitherState *d_state = nullptr; // [n_channels]
    std::vector<double> h_map;      // widened copy of the map as stored (float values for realsize 4)

    int init(int n_channels, int sample_rate, int realsize, int max_size, int max_samples_per_loop);
    void destroy();
};

} // namespace bfir
