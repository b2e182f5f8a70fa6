// Dispatch table over the per-size FFT translation units (fft_inst.cu).
#include "fft_dispatch.hpp"
#include <cstdlib>
#include <cstdint>

namespace bfir {

#define BFIR_DECL(tag, m)                                                                                      \
    cudaError_t launch_fwd_##tag##_m##m(int, dim3, cudaStream_t, const FwdArgs &, const void *, int, int);     \
    cudaError_t launch_inv_##tag##_m##m(int, dim3, cudaStream_t, const InvArgs &, const void *, int, int);     \
    cudaError_t launch_cfft_inv_##tag##_m##m(int, cudaStream_t, const CfftArgs &);
#define BFIR_FOR_F32(X) X(f32, 4) X(f32, 5) X(f32, 6) X(f32, 7) X(f32, 8) X(f32, 9) X(f32, 10) X(f32, 11) X(f32, 12) X(f32, 13) X(f32, 14)
#define BFIR_FOR_F64(X) X(f64, 4) X(f64, 5) X(f64, 6) X(f64, 7) X(f64, 8) X(f64, 9) X(f64, 10) X(f64, 11) X(f64, 12) X(f64, 13)
BFIR_FOR_F32(BFIR_DECL)
BFIR_FOR_F64(BFIR_DECL)
// 8 points per thread (double, per-CTA sizes 2^6 .. 2^12): no complex-FFT launcher in these units
#define BFIR_DECL_E8(tag, m)                                                                                   \
    cudaError_t launch_fwd_##tag##_m##m(int, dim3, cudaStream_t, const FwdArgs &, const void *, int, int);     \
    cudaError_t launch_inv_##tag##_m##m(int, dim3, cudaStream_t, const InvArgs &, const void *, int, int);
#define BFIR_FOR_F64E8(X) X(f64e8, 6) X(f64e8, 7) X(f64e8, 8) X(f64e8, 9) X(f64e8, 10) X(f64e8, 11) X(f64e8, 12)
BFIR_FOR_F64E8(BFIR_DECL_E8)
#define BFIR_FOR_F32E8(X) X(f32e8, 6) X(f32e8, 7) X(f32e8, 8) X(f32e8, 9) X(f32e8, 10) X(f32e8, 11) X(f32e8, 12) X(f32e8, 13)
BFIR_FOR_F32E8(BFIR_DECL_E8)
static const int kMinLog2M_e8 = 6, kMaxLog2M_e8 = 12, kMaxLog2M_e8_f32 = 13;

static const int kMinLog2M = 4, kMaxLog2M_f32 = 14, kMaxLog2M_f64 = 13;

#define BFIR_FWD_ENTRY(tag, m) launch_fwd_##tag##_m##m,
#define BFIR_INV_ENTRY(tag, m) launch_inv_##tag##_m##m,
static const fwd_launcher_t kFwdF32[] = { BFIR_FOR_F32(BFIR_FWD_ENTRY) };
static const inv_launcher_t kInvF32[] = { BFIR_FOR_F32(BFIR_INV_ENTRY) };
static const fwd_launcher_t kFwdF64[] = { BFIR_FOR_F64(BFIR_FWD_ENTRY) };
#define BFIR_CFFT_ENTRY(tag, m) launch_cfft_inv_##tag##_m##m,
static const cfft_launcher_t kCfftF32[] = { BFIR_FOR_F32(BFIR_CFFT_ENTRY) };
static const cfft_launcher_t kCfftF64[] = { BFIR_FOR_F64(BFIR_CFFT_ENTRY) };
static const inv_launcher_t kInvF64[] = { BFIR_FOR_F64(BFIR_INV_ENTRY) };
static const fwd_launcher_t kFwdF64E8[] = { BFIR_FOR_F64E8(BFIR_FWD_ENTRY) };
static const inv_launcher_t kInvF64E8[] = { BFIR_FOR_F64E8(BFIR_INV_ENTRY) };
static const fwd_launcher_t kFwdF32E8[] = { BFIR_FOR_F32E8(BFIR_FWD_ENTRY) };
static const inv_launcher_t kInvF32E8[] = { BFIR_FOR_F32E8(BFIR_INV_ENTRY) };

// 8 instead of 16 points per thread for a transform whose per-CTA size is 2^sub?
// Measured on B200 (tools/kernel_times.py, profiles/r01_fft_e8_vs_e16.jsonl). While a launch has at most one CTA
// per SM the kernels are a dependent-latency chain, and half the work per thread on twice the warps wins in both
// precisions (double cfg1 single stream: forward 22.6 -> 19.0 us, inverse 18.2 -> 16.7; 4 streams 30.3 -> 20.1 /
// 22.4 -> 18.2; product configuration 12.6 -> 10.2 / 10.8 -> 9.8; float cfg0 12.8 -> 11.1 / 11.4 -> 10.5; 128 float
// buffers of 1024 points 12.7 -> 10.3 / 11.6 -> 8.8). Larger launches are throughput-bound: 16 points per thread are
// as fast or faster (float cfg3 shape, 2048 buffers: 61.9 / 55.0 us against 71.5 / 59.9; double 4096 x 256:
// 22.8 / 21.6 against 23.1 / 24.0), except the double forward side up to 2048 points per CTA (16.9 -> 15.2 us).
// BFIR_FFT_E = 8 | 16 forces the answer where both exist.
static bool use_e8(int realsize, int sub, int r0, long long n_buffers, bool forward)
{
    if (r0 == 4) return false;
    if (sub < kMinLog2M_e8 || sub > (realsize == 8 ? kMaxLog2M_e8 : kMaxLog2M_e8_f32)) return false;
    static const int forced = [] { const char *env = getenv("BFIR_FFT_E"); return env ? atoi(env) : 0; }();
    if (forced == 8) return true;
    if (forced == 16) return false;
    static const int n_sm = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        return n;
    }();
    if (n_buffers * r0 <= n_sm) return true;
    return realsize == 8 && forward && sub <= 11;
}

static int max_sub(int realsize) { return realsize == 4 ? kMaxLog2M_f32 : kMaxLog2M_f64; }

bool rfft_supported(int realsize, int log2m)
{
    if (realsize != 4 && realsize != 8) return false;
    // one, two or (double precision, the largest size) four CTAs per transform
    return log2m >= kMinLog2M && log2m <= max_sub(realsize) + (realsize == 8 ? 2 : 1);
}

static int ilog2_r0(int r0) { return r0 == 4 ? 2 : (r0 == 2 ? 1 : 0); }

int rfft_choose_r0(int realsize, int log2m, long long n_buffers)
{
    if (log2m > max_sub(realsize) + 1) return 4;
    if (log2m > max_sub(realsize)) return 2;
    if (log2m - 1 < kMinLog2M) return 1;
    if (const char *env = getenv("BFIR_FFT_R0")) {
        const int v = atoi(env);
        if (v == 1 || v == 2) return v;
    }
    // measured on B200 (tools/kernel_times.py): with at most a few dozen buffers an 8192-point double
    // transform is latency-bound on one CTA per SM and two half-size CTAs are faster (cfg1 single stream:
    // forward 22.9 -> 21.7 us, inverse 22.6 -> 17.3 us); smaller or float transforms get slower (the
    // forward side reads every input twice), so they stay on one CTA unless the size requires two
    if (realsize == 8 && log2m >= 13 && n_buffers <= 64) return 2;
    return 1;
}

int cfft_max_log2m(int realsize) { return max_sub(realsize); }

cudaError_t launch_cfft_inverse(int realsize, int log2m, int batch, cudaStream_t stream, const CfftArgs &a)
{
    if ((realsize != 4 && realsize != 8) || log2m < kMinLog2M || log2m > max_sub(realsize) || batch < 1) return cudaErrorInvalidValue;
    return (realsize == 4 ? kCfftF32 : kCfftF64)[log2m - kMinLog2M](batch, stream, a);
}

// bulk-copy staging of the contiguous operands (rfft_kernels.cuh): needs 16-byte aligned sources and at least 16 bytes
static bool tma_enabled()
{
    static const bool on = [] { const char *env = getenv("BFIR_FFT_TMA"); return env ? atoi(env) != 0 : true; }();
    return on;
}

// De-interleaving through a thread-block cluster (rfft_kernels.cuh): Cc = the largest power of two <= 8 that divides the
// channels per stream, provided a 16-byte chunk of the (frame, channel)-ordered tile never straddles a frame (16 / bytes
// samples <= Cc) or the tile is the whole frame (Cc == C); launches whose channel range is not a whole number of
// clusters, sample sizes 1 and 3 and unaligned buffers keep the strided path. BFIR_FFT_CLUSTER=0 switches it off.
static int cluster_channels(int log2m, int r0, int fmt, int ch_per_stream, unsigned int grid_x, int ch_first, const void *raw, long long raw_stride_bytes)
{
    static const bool on = [] { const char *env = getenv("BFIR_FFT_CLUSTER"); return env ? atoi(env) != 0 : true; }();
    if (!on || r0 != 1 || log2m < 8 || raw == nullptr) return 0;
    const int bytes = fmt_bytes(fmt);
    if (bytes != 2 && bytes != 4 && bytes != 8) return 0;
    const int C = ch_per_stream, S = 16 / bytes;
    if (C < 2) return 0;
    // clusters of at most FOUR: 8-CTA clusters of one-CTA-per-SM kernels do not pack (measured: the 16 clusters of
    // cfg1 x 16 took two waves, 23 -> 35 us), 4-CTA clusters leave at most 16 of 148 SMs unused
    int cc = 4;
    while (cc > 1 && C % cc != 0) cc >>= 1;
    if (cc < 2) return 0;
    if (!(S <= cc || cc == C)) return 0;
    // worth it only where the strided path wastes at least three quarters of every sector it touches and the
    // cluster path gets four times more out of them (measured: stereo float, 50 % -> 100 %, is slower)
    {
        const int frame = C * bytes, run = cc * bytes;
        const double strided = (double)bytes / (frame < 32 ? frame : 32), tiled = (double)(run < 32 ? run : 32) / (frame < 32 ? frame : 32);
        if (tiled < 4.0 * strided) return 0;
    }
    if (cc != C && ((long long)C * bytes) % 16 != 0) return 0;      // frames of a partial tile must start on 16 bytes
    if (grid_x % (unsigned)cc != 0 || ch_first % cc != 0) return 0;
    if (((uintptr_t)raw & 15) != 0 || (raw_stride_bytes & 15) != 0) return 0;
    return cc;
}

cudaError_t launch_rfft_forward(int realsize, int log2m, int r0, dim3 grid, cudaStream_t stream, const FwdArgs &a0, const void *tw)
{
    if (!rfft_supported(realsize, log2m) || (r0 != 1 && r0 != 2 && r0 != 4)) return cudaErrorInvalidValue;
    const int sub = log2m - ilog2_r0(r0);
    if (sub < kMinLog2M || sub > max_sub(realsize)) return cudaErrorInvalidValue;
    FwdArgs a = a0;
    if (a.in_mode == IN_PLANAR2 && (grid.y < 1 || grid.y > 8)) return cudaErrorInvalidValue;
    a.tma = (tma_enabled() && r0 == 1 && a.in_mode == IN_RAW_PREV && a.prev != nullptr && ((uintptr_t)a.prev & 15) == 0
             && (((size_t)realsize << log2m) & 15) == 0) ? 1 : 0;
    a.cluster = 0;
    if (a.in_mode == IN_RAW_PREV && a.prev != nullptr && grid.y == 1)
        a.cluster = cluster_channels(log2m, r0, a.fmt, a.ch_per_stream, grid.x, a.ch_base % a.ch_per_stream, a.in, a.in_stride_x);
    // table length N = 2 * 2^log2m: shift for the sub-transform twiddles, 0 for the W_N^k of the split step
    if (use_e8(realsize, sub, r0, (long long)grid.x * grid.y, true)) return (realsize == 4 ? kFwdF32E8 : kFwdF64E8)[sub - kMinLog2M_e8](r0, grid, stream, a, tw, 1 + ilog2_r0(r0), 0);
    return (realsize == 4 ? kFwdF32 : kFwdF64)[sub - kMinLog2M](r0, grid, stream, a, tw, 1 + ilog2_r0(r0), 0);
}

cudaError_t launch_rfft_inverse(int realsize, int log2m, int r0, dim3 grid, cudaStream_t stream, const InvArgs &a0, const void *tw)
{
    if (!rfft_supported(realsize, log2m) || (r0 != 1 && r0 != 2 && r0 != 4)) return cudaErrorInvalidValue;
    const int sub = log2m - ilog2_r0(r0);
    if (sub < kMinLog2M || sub > max_sub(realsize)) return cudaErrorInvalidValue;
    InvArgs a = a0;
    if (a.n_multi > 0) {   // several blocks in one launch: grid.y = blocks
        if (a.n_multi > 8 || a.head_x != nullptr) return cudaErrorInvalidValue;
        grid.y = (unsigned)a.n_multi;
    }
    uintptr_t in_bits = (uintptr_t)a.in, out_bits = (uintptr_t)a.out;
    if (a.n_multi > 0) { in_bits = out_bits = 0; for (int k = 0; k < a.n_multi; k++) { in_bits |= (uintptr_t)a.in_multi[k]; out_bits |= (uintptr_t)a.out_multi[k]; } }
    a.tma = (tma_enabled() && r0 == 1 && a.head_x == nullptr && (in_bits & 15) == 0
             && (((size_t)a.in_stride_x * realsize) & 15) == 0) ? 1 : 0;
    a.cluster = 0;
    if (a.out_mode == OUT_RAW && (grid.y == 1 || a.n_multi > 0))
        a.cluster = cluster_channels(log2m, r0, a.fmt, a.ch_per_stream, grid.x, (a.ch_base - a.raw_ch_base) % a.ch_per_stream, (const void *)out_bits, a.out_stride_x);
    if (use_e8(realsize, sub, r0, (long long)grid.x * grid.y, false)) return (realsize == 4 ? kInvF32E8 : kInvF64E8)[sub - kMinLog2M_e8](r0, grid, stream, a, tw, 1 + ilog2_r0(r0), 0);
    return (realsize == 4 ? kInvF32 : kInvF64)[sub - kMinLog2M](r0, grid, stream, a, tw, 1 + ilog2_r0(r0), 0);
}

} // namespace bfir
