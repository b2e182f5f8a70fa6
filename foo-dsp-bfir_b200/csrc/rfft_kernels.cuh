// Real <-> half-complex transforms of one convolver buffer per CTA, with the surrounding stages of
// the reference's block loop fused into the load and store phases:
//
//   forward kernel  = raw2cbuf (raw2real + [prev|cur] framing)         brutefir.cpp:255-260, fftw_convolver.cpp:157-185
//                   + time2freq (FFTW_R2HC)                            brutefir.cpp:263,     fftw_convolver.cpp:188-212
//                   + mixnscale(MIXMODE_INPUT, n_bufs=1): scale, HC->ORD brutefir.cpp:273-277, fftw_convolver.cpp:883-907
//                   also coeffs2cbuf / runtime_coeffs2cbuf             fftw_convolver.cpp:475-567
//   inverse kernel  = mixnscale(MIXMODE_OUTPUT, n_bufs=1): scale, ORD->HC brutefir.cpp:303-307, fftw_convolver.cpp:1163-1186
//                   + freq2time (FFTW_HC2R)                            brutefir.cpp:311,     fftw_convolver.cpp:351-375
//                   + NaN/Inf probe on sample 0                        brutefir.cpp:316-321
//                   + cbuf2raw (real2raw, no-dither paths + statistics) brutefir.cpp:326-331, fftw_convolver.cpp:406-466
//
// Buffer layouts: T = [prev|cur] time, HC = FFTW half-complex, ORD = groups of
// 8 reals [Re k..k+3 | Im k..k+3] with Re X_{N/2} in slot 4.
#pragma once
#include "fft_core.cuh"
#include "codec.cuh"
#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

namespace bfir {

enum { LAYOUT_ORD = 0, LAYOUT_HC = 1 };
enum {
    IN_TIME = 0,     // N reals, T layout (time2freq)
    IN_RAW_PREV = 1, // engine: interleaved raw block + per-channel previous block
    IN_COEFF = 2,    // planar coefficients, partition blockIdx.y: [0_L | h[y*L .. ) * scale] (coeffs2cbuf)
    IN_UPPER = 3,    // [0_L | src[0..L)] (runtime_coeffs2cbuf)
    IN_PLANAR2 = 4   // engine, several blocks per launch (blockIdx.y = block): [planar previous block | planar current block] of
                     // channel bx, rows lo_multi[y] + bx L and hi_multi[y] + bx L (de-interleaved beforehand by raw_to_planar_kernel)
};
enum {
    OUT_TIME = 0,    // all N reals (freq2time)
    OUT_RAW = 1,     // engine: first L samples -> interleaved raw (float formats / integer without dither)
    OUT_REAL_L = 2   // engine: first L samples -> planar real scratch (consumed by the dither kernel)
};

// Fused partition-shard reduce ("fused variant"): instead of writing its PARTIAL spectrum
// of reduced channel `ch` to local memory for a later NCCL reduce, the producing kernel stores it
// straight into the receive buffer of the rank that owns `ch` -- a peer-mapped pointer, i.e. plain
// st.global over NVLink -- so the transfer overlaps the partition sum tile by tile. Receive buffer
// of rank q: [BFIR_PEER_PHASES][world (source rank)][cpr (owned channels)][N], then one arrival flag per source
// rank. Phases 0 / 1: block parity of the one-block calls; phases 2 .. 17: (call parity, block) of the four- and eight-block calls
// (phase 2 + 8 parity + b; consecutive calls differ in parity).
#define BFIR_MAX_PEERS 8
#define BFIR_PEER_PHASES 18
struct PeerPush {
    void *recv[BFIR_MAX_PEERS];
    int world, self, cpr, enabled;
    long long flag_offset;      // bytes from the start of a receive buffer to its unsigned int flags[world]
    // sharded input stage (crossbar engines, four-block staged calls): every rank transforms its own in_cpr inputs and
    // stores the spectra into every peer's input region [2 call parities * 4 blocks][n_inputs][N]; second flag array
    long long xin_offset;       // bytes from the start of a receive buffer to that region (0: not allocated)
    long long flag_in_offset;   // ... and to unsigned int flags_in[world]
    int in_cpr, n_inputs;
};
template <class T> BFIR_HD T *peer_dst(const PeerPush &p, int ch, int N, unsigned int phase)
{
    const int q = ch / p.cpr, local = ch - q * p.cpr;
    return (T *)p.recv[q] + (((long long)phase * p.world + p.self) * p.cpr + local) * N;
}

// device-resident engine state: lets one CUDA graph be replayed for every block
struct EngineState {
    unsigned int blockcounter;  // brutefir.hpp:106
    int first_bad_channel;      // lowest channel whose output sample 0 was NaN/Inf this block
    unsigned int cur_slot;      // delay-line slot of the block in flight, left by the forward kernel for a fused head term
};

struct FwdArgs {
    int in_mode, out_layout;
    const void *in;          // IN_TIME / IN_UPPER: reals; IN_RAW_PREV: raw bytes; IN_COEFF: planar coefficients
    long long in_stride_x;   // per blockIdx.x (channel / buffer), in elements (bytes for raw: per stream)
    long long in_stride_y;   // per blockIdx.y
    void *out;
    long long out_stride_x, out_stride_y; // elements
    double scale_in;         // applied to the samples (coefficient scale)
    double scale_out;        // applied to the spectrum (input scale, or 1/N for coefficients)
    // IN_RAW_PREV
    void *prev;              // [2][n_channels][L] previous block, ping-pong by blockcounter parity
    int fmt, ch_per_stream, n_channels;
    int ch_base;             // first channel of this launch (channel-group pipelining): bx = blockIdx.x + ch_base
    unsigned int prev_parity; // blockcounter & 1, tracked by the host so that no load depends on the state word
    const EngineState *state; // out slot = (blockcounter + slot_offset) % n_slots when state != NULL
    int n_slots;             // delay-line slots per channel (the engine keeps one more than partitions)
    int n_parts;             // filter partitions: procblocks counts up to this (brutefir.cpp:265-268)
    int slot_offset;         // 1: the second block of a pair, transformed before the first one has been counted
    int use_abs_block;       // 1: the block index is abs_block, given by the host (stage pipeline: the device counter lags behind)
    unsigned int abs_block;
    int *procblocks;         // [channels], brutefir.cpp:265-268
    unsigned char *pb_inc;   // [channels], 1 when procblocks was incremented by this launch
    // IN_COEFF
    int coeff_len;           // valid coefficients per channel
    int *nonfinite;          // set to 1 when a scaled coefficient is NaN/Inf (fftw_convolver.cpp:493-497)
    int tma;                 // 1: IN_RAW_PREV, one CTA per transform: the previous block arrives by bulk copy (set by launch_rfft_forward)
    int cluster;             // > 1: IN_RAW_PREV, the CTAs of `cluster` neighbouring channels of a stream form a thread-block cluster
                             // and de-interleave the raw block together (set by launch_rfft_forward; see fwd_load_cluster)
    const void *lo_multi[8], *hi_multi[8];   // IN_PLANAR2: planar [channels][L] rows of the previous / current block of block y
};

// index of the block a forward launch transforms
template <class Dummy> BFIR_HD unsigned int fwd_block(const FwdArgs &a)
{
    return a.use_abs_block ? a.abs_block : a.state->blockcounter + (unsigned int)a.slot_offset;
}

struct InvArgs {
    int in_layout, out_mode;
    const void *in;
    long long in_stride_x;   // elements
    double scale_in;         // applied to the spectrum (output scale)
    void *out;               // OUT_TIME / OUT_REAL_L: reals; OUT_RAW: raw bytes
    long long out_stride_x;  // elements (bytes per stream for OUT_RAW)
    int fmt, ch_per_stream;
    int ch_base;             // first channel of this launch
    int raw_ch_base;         // subtracted from the channel for raw addressing (compact own-channel output)
    double ovf_max;          // bfoverflow_t.max
    OverflowStats *stats;    // [channels]
    EngineState *state;      // probe + blockcounter++ (engine only)
    int *host_flag;          // mapped pinned host word, set to 1 on a non-finite probe (lets the host skip a D2H)
    // look-ahead partition sum: `in` holds sum_{i>=1} X[t-i] H[i], computed before the block arrived; the load
    // phase adds the head term X[t] H[0] (both ORD, slot state->cur_slot of the delay line) on the fly
    const void *head_x;      // delay line [channels][head_slots][N]; NULL = no head term
    const void *head_h;      // coefficients [channels][..][N], partition 0 at the start of each channel
    long long head_x_stride, head_h_stride; // elements per channel
    const int *head_blocks;  // [channels] coefficient partitions loaded (0: the channel has no filter)
    const int *head_map;     // [channels] coefficient set of each channel, NULL = its own (bfir_set_coeff_map)
    int tma;                 // 1: one CTA per transform, no head term: the input spectrum arrives by bulk copy (set by launch_rfft_inverse)
    int cluster;             // > 1: OUT_RAW, the CTAs of `cluster` neighbouring channels of a stream form a thread-block cluster and
                             // interleave their output together (set by launch_rfft_inverse; see inv_store_cluster)
    // several consecutive blocks in ONE launch (grid.y = n_multi, the stage pipeline's output stage): block y reads
    // in_multi[y] and writes out_multi[y]; `in` / `out` are ignored then
    int n_multi;
    const void *in_multi[8];
    void *out_multi[8];
};

// input / output of the block a CTA works on (blockIdx.y selects it in a multi-block launch)
BFIR_HD const void *inv_in(const InvArgs &a)
{
#ifdef __CUDA_ARCH__
    if (a.n_multi > 0) return a.in_multi[blockIdx.y];
#endif
    return a.in;
}
BFIR_HD void *inv_out(const InvArgs &a)
{
#ifdef __CUDA_ARCH__
    if (a.n_multi > 0) return a.out_multi[blockIdx.y];
#endif
    return a.out;
}

template <class T> BFIR_HD T tw_re(const cpx<T> &w) { return w.x; }

// ------------------------------------------------------------------------------------------------
// spectrum access by bin for both layouts. N = 2M. X_0 and X_M are real.
template <class T> BFIR_HD cpx<T> spec_load(const T *s, int layout, int k, int M)
{
    cpx<T> r;
    if (layout == LAYOUT_ORD) {
        const int base = ((k >> 2) << 3) + (k & 3);
        if (k == 0) { r.x = s[0]; r.y = (T)0; }
        else if (k == M) { r.x = s[4]; r.y = (T)0; }
        else { r.x = s[base]; r.y = s[base + 4]; }
    } else {
        if (k == 0) { r.x = s[0]; r.y = (T)0; }
        else if (k == M) { r.x = s[M]; r.y = (T)0; }
        else { r.x = s[k]; r.y = s[2 * M - k]; }
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// CTA-level split of one transform (R0 = 1 or 2 CTAs per buffer, blockIdx.z = r):
// a radix-R0 decimation-in-frequency pre-pass is folded into the load phase, after which CTA r owns an
// independent Ms = M/R0 point transform:
//     forward:  s_r[n] = W_M^(n r) * sum_j z[n + j Ms] W_R0^(j r)      ->  Z[R0 k + r] = FFT_Ms(s_r)[k]
//     inverse:  s_r[k] = W_M^(-k r) * sum_j Z'[k + j Ms] W_R0^(-j r)   ->  z[R0 n + r] = IFFT_Ms(s_r)[n]
// For R0 = 2 the real-FFT split partner of bin k = 2k'+r is M-k = 2(Ms-k')  (r = 0) or 2(Ms-1-k')+1
// (r = 1): the same CTA, so no exchange between the two CTAs is ever needed. R0 = 2 doubles the
// largest block length (L = 32768 float, 16384 double) and halves the per-CTA latency of big transforms.

// R0 = 4 (four CTAs per transform, double-precision 65536-point real transforms: 32768 complex points do not fit
// two CTAs' shared memory): the radix-4 decimation-in-frequency pre-pass of residue r on four points Ms apart,
//     forward: y_r = z0 + z1 W4^r + z2 W4^2r + z3 W4^3r, W4 = -i;   inverse: the conjugate roots.
// The real-FFT split partner of bin k = 4k'+r is M-k = 4(Ms-k') (r = 0) or 4(Ms-1-k') + (4-r): residues 1 and 3
// pair ACROSS CTAs, so the forward kernel's four CTAs form a thread-block cluster and read the partner's
// sub-transform through distributed shared memory.
template <bool INV, class C> BFIR_HD C dif4_residue(int r, C z0, C z1, C z2, C z3)
{
    if (r == 0) return cadd(cadd(z0, z2), cadd(z1, z3));
    if (r == 2) return csub(cadd(z0, z2), cadd(z1, z3));
    const C d = mul_mi<INV>(csub(z1, z3));            // forward: -i (z1 - z3); inverse: +i (z1 - z3)
    return r == 1 ? cadd(csub(z0, z2), d) : csub(csub(z0, z2), d);
}
// w^r for r = 0..3
template <class C> BFIR_HD C cpow_small(C w, int r)
{
    C one; one.x = 1; one.y = 0;
    if (r == 0) return one;
    if (r == 1) return w;
    const C w2 = cmul(w, w);
    return r == 2 ? w2 : cmul(w2, w);
}
template <int R0> struct log2_r0 { static constexpr int value = R0 == 4 ? 2 : (R0 == 2 ? 1 : 0); };

// per-thread constants of the input access, computed once (the state word must not be re-read per element:
// the stores into the previous-block buffer would force the compiler to reload it every time)
template <class T> struct FwdCtx {
    const cpx<T> *plain;     // IN_TIME / IN_UPPER
    const T *coeff;          // IN_COEFF
    const cpx<T> *prev_rd;   // IN_RAW_PREV: previous block (read), current block (write)
    cpx<T> *prev_wr;
    const uint8_t *raw;
    long long step;
    T sc;
};

template <class T, int LOG2M>
BFIR_HD FwdCtx<T> fwd_ctx(int bx, int by, const FwdArgs &a)
{
    constexpr int L = 1 << LOG2M;
    typedef cpx<T> C;
    FwdCtx<T> c;
    c.plain = (const C *)((const T *)a.in + bx * a.in_stride_x + by * a.in_stride_y);
    c.coeff = (const T *)a.in + bx * a.in_stride_x;
    c.sc = (T)a.scale_in;
    c.prev_rd = NULL; c.prev_wr = NULL; c.raw = NULL; c.step = 0;
    if (a.in_mode == IN_PLANAR2) {
        c.prev_rd = (const C *)((const T *)a.lo_multi[by] + (long long)bx * L);
        c.plain = (const C *)((const T *)a.hi_multi[by] + (long long)bx * L);
    }
    if (a.in_mode == IN_RAW_PREV) {
        // [previous block | current block], previous kept in a ping-pong pair that follows the reference's
        // input_timecbuf[n][curbuf] (brutefir.cpp:255-260, 337)
        const unsigned int par = a.prev_parity & 1u;
        c.prev_rd = (const C *)((const T *)a.prev + ((long long)par * a.n_channels + bx) * L);
        c.prev_wr = (C *)((T *)a.prev + ((long long)(par ^ 1u) * a.n_channels + bx) * L);
        const int stream = bx / a.ch_per_stream, ch = bx - stream * a.ch_per_stream;
        const int bytes = fmt_bytes(a.fmt);
        c.step = (long long)a.ch_per_stream * bytes;
        c.raw = (const uint8_t *)a.in + (long long)stream * a.in_stride_x + (long long)ch * bytes;
    }
    return c;
}

// element z[m] of the packed input, m in [0, M); UPPER tells at compile time that m >= M/2
template <class T, int LOG2M, bool UPPER>
BFIR_HD cpx<T> fwd_elem(int m, int by, int r, const FwdArgs &a, const FwdCtx<T> &c, bool &bad)
{
    constexpr int M = 1 << LOG2M, L = M;
    typedef cpx<T> C;
    if (a.in_mode == IN_TIME) {
        return c.plain[m];
    } else if (a.in_mode == IN_PLANAR2) {
        if (!UPPER) return c.prev_rd[m];
        return c.plain[m - M / 2];
    } else if (a.in_mode == IN_UPPER) {
        if (!UPPER) return mk<T>((T)0, (T)0);
        return c.plain[m - M / 2];
    } else if (a.in_mode == IN_COEFF) {
        if (!UPPER) return mk<T>((T)0, (T)0);
        const long long g = (long long)by * L + 2 * (m - M / 2);   // index into the channel's coefficients
        T c0 = (T)0, c1 = (T)0;
        if (g < a.coeff_len) c0 = c.coeff[g] * c.sc;
        if (g + 1 < a.coeff_len) c1 = c.coeff[g + 1] * c.sc;
        bad = bad || !(c0 - c0 == (T)0) || !(c1 - c1 == (T)0);     // NaN or Inf
        return mk<T>(c0, c1);
    } else { // IN_RAW_PREV
        if (!UPPER) return c.prev_rd[m];
        const int n = m - M / 2;                                   // complex index inside the current block
        const uint8_t *p = c.raw + (long long)(2 * n) * c.step;
        const C z = mk<T>(load_raw<T>(p, a.fmt), load_raw<T>(p + c.step, a.fmt));
        if (r == 0) c.prev_wr[n] = z;
        return z;
    }
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Bulk-asynchronous staging (TMA, cp.async.bulk + mbarrier) of the CONTIGUOUS operands of a transform into the CTA's
// shared-memory buffer, which is idle until the first pass stores into it: the inverse kernel's input spectrum
// (N reals of one channel) and the forward kernel's previous block (L reals). One elected thread arms the barrier
// with the byte count and issues the copies; the TMA unit streams the lines while the threads compute their twiddles
// and issue the loads that cannot be bulk copies (the interleaved raw block); everybody then waits on the barrier's
// phase. Replaces 16-32 LDG.128 per thread that each had to be tracked by the LSU.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "BFIR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra BFIR_DONE;\n"
                 "bra BFIR_WAIT;\n"
                 "BFIR_DONE:\n"
                 "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one thread: `bytes` (a multiple of 16) from global to shared memory in pieces of at most 32 KB
__device__ __forceinline__ void bulk_stage(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    mbar_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u)
        bulk_g2s((char *)dst + off, (const char *)src + off, min(32768u, bytes - off), bar);
}

#endif

// The sample format is a run-time engine parameter but must be a compile-time constant inside the unrolled load /
// store loops: with the switch inside load_raw() every sample's load sat behind its own branch, and the loads of
// one thread were issued one by one, each waiting out a full memory latency (ncu source page, round 1: 40 % of the
// forward kernel's stall samples). BFIR_FMT_SWITCH hoists the switch around the whole loop.
#define BFIR_FMT_CASE(f, ...) case f: { constexpr int FMT = f; __VA_ARGS__; } break;
#ifdef BFIR_LEAN_FORMATS   // experiment: only the two little-endian float formats (code size of the transform kernels)
#define BFIR_FMT_SWITCH(fmt, ...)                                                                        \
    switch (fmt) {                                                                                       \
        BFIR_FMT_CASE(FMT_FLOAT_LE, __VA_ARGS__) BFIR_FMT_CASE(FMT_FLOAT64_LE, __VA_ARGS__)               \
        default: break;                                                                                  \
    }
#else
#define BFIR_FMT_SWITCH(fmt, ...)                                                                        \
    switch (fmt) {                                                                                       \
        BFIR_FMT_CASE(FMT_S8, __VA_ARGS__) BFIR_FMT_CASE(FMT_S16_LE, __VA_ARGS__) BFIR_FMT_CASE(FMT_S16_BE, __VA_ARGS__) \
        BFIR_FMT_CASE(FMT_S24_LE, __VA_ARGS__) BFIR_FMT_CASE(FMT_S24_BE, __VA_ARGS__) BFIR_FMT_CASE(FMT_S32_LE, __VA_ARGS__) \
        BFIR_FMT_CASE(FMT_S32_BE, __VA_ARGS__) BFIR_FMT_CASE(FMT_FLOAT_LE, __VA_ARGS__) BFIR_FMT_CASE(FMT_FLOAT_BE, __VA_ARGS__) \
        BFIR_FMT_CASE(FMT_FLOAT64_LE, __VA_ARGS__) BFIR_FMT_CASE(FMT_FLOAT64_BE, __VA_ARGS__)               \
        default: break;                                                                                  \
    }
#endif

// engine input (IN_RAW_PREV) for one thread, sample format known at compile time: the raw samples of the current
// block (complex index n = t + j NTs of the block, i.e. frames 2n and 2n+1) and the previous block from its planar
// ping-pong buffer; all loads are independent and issued back to back
template <class T, int LOG2MS, int R0, int LOG2E, int FMT, bool STAGED = false>
BFIR_HD void fwd_load_raw_prev(int t, int r, cpx<T> (&v)[1 << LOG2E], const cpx<T> wpre, const FwdCtx<T> &c, const cpx<T> *staged_prev = NULL, void *bar = NULL)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E;
    constexpr int NH = R0 == 1 ? E / 2 : E;       // complex points of the current block per thread
    constexpr int CH = R0 == 2 ? 4 : NH;          // complex loads in flight per batch and source (R0 = 2: register budget)
    typedef cpx<T> C;
    if constexpr (STAGED) {
        // R0 = 1: the raw samples first (their loads are in flight while the bulk copy of the previous block lands),
        // then the previous block out of shared memory
        C hi[NH];
#pragma unroll
        for (int j = 0; j < NH; j++) {
            const uint8_t *p = c.raw + (long long)(2 * (t + j * NT)) * c.step;
            hi[j] = mk<T>(load_raw<T>(p, FMT), load_raw<T>(p + c.step, FMT));
        }
#ifdef __CUDA_ARCH__
        mbar_wait((uint64_t *)bar, 0);
#endif
#pragma unroll
        for (int j = 0; j < NH; j++) v[j] = staged_prev[t + j * NT];
#pragma unroll
        for (int j = 0; j < NH; j++) { c.prev_wr[t + j * NT] = hi[j]; v[j + NH] = hi[j]; }
    } else if constexpr (R0 == 4) {
        // n = t + j NT in [0, Ms): previous block at complex indices n and n + Ms, current block likewise (M/2 = 2 Ms)
#pragma unroll
        for (int j0 = 0; j0 < E; j0 += 2) {
            C lo0[2], lo1[2], hi0[2], hi1[2];
#pragma unroll
            for (int j = 0; j < 2; j++) { lo0[j] = c.prev_rd[t + (j0 + j) * NT]; lo1[j] = c.prev_rd[t + (j0 + j) * NT + MS]; }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint8_t *p = c.raw + (long long)(2 * (t + (j0 + j) * NT)) * c.step;
                const uint8_t *q = c.raw + (long long)(2 * (t + (j0 + j) * NT + MS)) * c.step;
                hi0[j] = mk<T>(load_raw<T>(p, FMT), load_raw<T>(p + c.step, FMT));
                hi1[j] = mk<T>(load_raw<T>(q, FMT), load_raw<T>(q + c.step, FMT));
            }
            if (r == 0) {
#pragma unroll
                for (int j = 0; j < 2; j++) { c.prev_wr[t + (j0 + j) * NT] = hi0[j]; c.prev_wr[t + (j0 + j) * NT + MS] = hi1[j]; }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const C y = dif4_residue<false>(r, lo0[j], lo1[j], hi0[j], hi1[j]);
                v[j0 + j] = r == 0 ? y : cmul(y, cpow_small(cmul(wpre, thread_root<T, 4 * E>(j0 + j)), r));
            }
        }
    } else {
#pragma unroll
    for (int j0 = 0; j0 < NH; j0 += CH) {
        C lo[CH], hi[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) lo[j] = c.prev_rd[t + (j0 + j) * NT];
#pragma unroll
        for (int j = 0; j < CH; j++) {
            const uint8_t *p = c.raw + (long long)(2 * (t + (j0 + j) * NT)) * c.step;
            hi[j] = mk<T>(load_raw<T>(p, FMT), load_raw<T>(p + c.step, FMT));
        }
        if (r == 0) {
#pragma unroll
            for (int j = 0; j < CH; j++) c.prev_wr[t + (j0 + j) * NT] = hi[j];
        }
        if (R0 == 1) {
#pragma unroll
            for (int j = 0; j < CH; j++) { v[j0 + j] = lo[j]; v[j0 + j + NH] = hi[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < CH; j++) {
                if (r == 0) v[j0 + j] = cadd(lo[j], hi[j]);
                else v[j0 + j] = cmul(csub(lo[j], hi[j]), cmul(wpre, thread_root<T, 2 * E>(j0 + j)));   // W_M^n = W_N^(2n)
            }
        }
    }
    }
}

// engine bookkeeping of a forward launch: the slot of the block in flight, procblocks (brutefir.cpp:265-268)
template <class Dummy> BFIR_HD void fwd_bookkeeping(int t, int bx, int by, int r, const FwdArgs &a)
{
    if (a.in_mode == IN_RAW_PREV && a.state != NULL && t == 0 && r == 0 && by == 0 && bx == a.ch_base)
        const_cast<EngineState *>(a.state)->cur_slot = fwd_block<int>(a) % (unsigned int)a.n_slots;
    if (a.in_mode == IN_RAW_PREV && t == 0 && r == 0 && a.procblocks != NULL) {
        const int pb = a.procblocks[bx];
        const bool inc = pb < a.n_parts;
        if (inc) a.procblocks[bx] = pb + 1;
        a.pb_inc[bx] = inc ? 1 : 0;
    }
}

// forward, phase 0: thread t builds s_r[n], n = t + i*NTs
template <class T, int LOG2MS, int R0, int LOG2E = 4, bool STAGED = false>
BFIR_HD void fwd_load(int t, int bx, int by, int r, cpx<T> (&v)[1 << LOG2E], const cpx<T> *__restrict__ tw, int tw_shift_n, const FwdArgs &a,
                      const cpx<T> *staged_prev = NULL, void *bar = NULL)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E, LOG2M = LOG2MS + log2_r0<R0>::value;
    typedef cpx<T> C;
    bool bad = false;
    cpx<T> wpre = mk<T>((T)1, (T)0);
    if (R0 == 2 && r == 1) wpre = tw[(2 * t) << tw_shift_n];     // W_M^t = W_N^(2t); W_M^(t + i NTs) = W_M^t * root32(i)
    if (R0 == 4 && r != 0) wpre = tw[(2 * t) << tw_shift_n];     // W_M^t; W_M^(t + i NTs) = W_M^t * root64(i), then the r-th power
    const FwdCtx<T> ctx = fwd_ctx<T, LOG2M>(bx, by, a);
    if (a.in_mode == IN_RAW_PREV) {
        BFIR_FMT_SWITCH(a.fmt, (fwd_load_raw_prev<T, LOG2MS, R0, LOG2E, FMT, STAGED>(t, r, v, wpre, ctx, staged_prev, bar)))
    } else if (R0 == 1) {
#pragma unroll
        for (int i = 0; i < E / 2; i++) v[i] = fwd_elem<T, LOG2M, false>(t + i * NT, by, r, a, ctx, bad);
#pragma unroll
        for (int i = E / 2; i < E; i++) v[i] = fwd_elem<T, LOG2M, true>(t + i * NT, by, r, a, ctx, bad);
    } else if (R0 == 4) {
#pragma unroll
        for (int i = 0; i < E; i++) {
            const int n = t + i * NT;
            const C z0 = fwd_elem<T, LOG2M, false>(n, by, r, a, ctx, bad);
            const C z1 = fwd_elem<T, LOG2M, false>(n + MS, by, r, a, ctx, bad);
            const C z2 = fwd_elem<T, LOG2M, true>(n + 2 * MS, by, r, a, ctx, bad);
            const C z3 = fwd_elem<T, LOG2M, true>(n + 3 * MS, by, r, a, ctx, bad);
            const C y = dif4_residue<false>(r, z0, z1, z2, z3);
            v[i] = r == 0 ? y : cmul(y, cpow_small(cmul(wpre, thread_root<T, 4 * E>(i)), r));
        }
    } else {
#pragma unroll
        for (int i = 0; i < E; i++) {
            const int n = t + i * NT;
            const C lo = fwd_elem<T, LOG2M, false>(n, by, r, a, ctx, bad);
            const C hi = fwd_elem<T, LOG2M, true>(n + MS, by, r, a, ctx, bad);
            if (r == 0) v[i] = cadd(lo, hi);
            else v[i] = cmul(csub(lo, hi), cmul(wpre, thread_root<T, 2 * E>(i)));   // W_M^n = W_N^(2n)
        }
    }
    if (a.in_mode == IN_COEFF && bad) *a.nonfinite = 1;
    fwd_bookkeeping<int>(t, bx, by, r, a);
}

// forward, phase 2: sub-transform result (natural order, padded smem) -> X_k, k = R0 k' + r, scaled,
// stored in ORD or HC layout
// `partner`: where Z_{M-k} lives -- this CTA's own buffer except for R0 = 4, r = 1 / 3 (the other one's, through DSMEM)
template <class T, int LOG2MS, int R0, int LOG2E = 4>
BFIR_HD void fwd_split_store(int t, int bx, int by, int r, const cpx<T> *smem, const cpx<T> *__restrict__ tw, int tw_shift_n, const FwdArgs &a,
                             const cpx<T> *partner = NULL)
{
    if (partner == NULL) partner = smem;
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E, M = MS * R0, N = 2 * M;
    typedef cpx<T> C;
    long long off = bx * a.out_stride_x + by * a.out_stride_y;
    if (a.state != NULL) off += (long long)(fwd_block<int>(a) % (unsigned int)a.n_slots) * a.out_stride_y;
    T *out = (T *)a.out + off;
    const T sc = (T)a.scale_out;
    const C wbase = tw[(R0 * t + r) << tw_shift_n];   // W_N^k of i = 0; the thread's bins are N/(2E) apart
#pragma unroll
    for (int i = 0; i < E; i++) {
        const int kp = t + i * NT;
        const int k = R0 * kp + r;
        const C zk = smem[fft_pad(kp)];
        if (k == 0) {
            const T dc = (zk.x + zk.y) * sc, ny = (zk.x - zk.y) * sc;
            out[0] = dc;
            if (a.out_layout == LAYOUT_ORD) out[4] = ny; else out[M] = ny;
        } else {
            const C zm = partner[fft_pad(r == 0 ? MS - kp : MS - 1 - kp)];   // Z_{M-k}
            // E = (Z_k + conj Z_{M-k})/2, O = (Z_k - conj Z_{M-k})/(2i), X_k = E + W_N^k O
            const T er = (T)0.5 * (zk.x + zm.x), ei = (T)0.5 * (zk.y - zm.y);
            const T dr = (T)0.5 * (zk.x - zm.x), di = (T)0.5 * (zk.y + zm.y);
            const T o_r = di, o_i = -dr;
            const C w = cmul(wbase, thread_root<T, 2 * E>(i));
            const T xr = (er + (w.x * o_r - w.y * o_i)) * sc;
            const T xi = (ei + (w.x * o_i + w.y * o_r)) * sc;
            if (a.out_layout == LAYOUT_ORD) {
                const int base = ((k >> 2) << 3) + (k & 3);
                out[base] = xr;
                out[base + 4] = xi;
            } else {
                out[k] = xr;
                out[N - k] = xi;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Z'_k = (X_k + conj X_{M-k}) + i conj(W_N^k) (X_k - conj X_{M-k}), the packed spectrum whose inverse
// complex transform is z[n] = x[2n] + i x[2n+1]
// bin k of `in`, plus the head term X_k H_k when hx != NULL (the real bins 0 and M come back with a zero
// imaginary part from spec_load, so the complex product is the reference's d1s/d2s real product there)
template <class T>
BFIR_HD cpx<T> spec_load_head(const T *in, int layout, int k, int M, const T *hx, const T *hh)
{
    cpx<T> s = spec_load<T>(in, layout, k, M);
    if (hx != NULL) {
        const cpx<T> x = spec_load<T>(hx, LAYOUT_ORD, k, M), h = spec_load<T>(hh, LAYOUT_ORD, k, M);
        s.x += x.x * h.x - x.y * h.y;
        s.y += x.x * h.y + x.y * h.x;
    }
    return s;
}

template <class T>
BFIR_HD cpx<T> inv_elem(const T *in, int layout, int k, int M, T sc, const cpx<T> w, const T *hx = NULL, const T *hh = NULL)
{
    typedef cpx<T> C;
    C xk = spec_load_head<T>(in, layout, k, M, hx, hh);
    C xm = spec_load_head<T>(in, layout, M - k, M, hx, hh);
    xk.x *= sc; xk.y *= sc; xm.x *= sc; xm.y *= sc;
    const T er = xk.x + xm.x, ei = xk.y - xm.y;   // X_k + conj X_{M-k}
    const T dr = xk.x - xm.x, di = xk.y + xm.y;   // X_k - conj X_{M-k}
    const T pr = w.x * dr + w.y * di;             // conj(w) * d, w = W_N^k
    const T pi = w.x * di - w.y * dr;
    return mk<T>(er - pi, ei + pr);               // e + i p
}

// inverse, phase 0: thread t builds s_r[k], k = t + i*NTs
template <class T, int LOG2MS, int R0, bool HEAD, int LOG2E = 4>
BFIR_HD void inv_load_impl(int t, int r, cpx<T> (&v)[1 << LOG2E], const cpx<T> *__restrict__ tw, int tw_shift_n, const T *in, int layout, T sc,
                           const T *hx, const T *hh)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E, M = MS * R0;
    typedef cpx<T> C;
    // one table look-up per thread: the thread's bins k = t + i NT are N/(2E) (one CTA) or N/(4E) (two CTAs) apart
    const C wbase = tw[t << tw_shift_n];                                   // W_N^t
    C wpre = mk<T>((T)1, (T)0);
    if (R0 == 2 && r == 1) wpre = tw[(2 * t) << tw_shift_n];               // W_M^t
    if constexpr (R0 == 4) {
        // Z' at k + j Ms, j = 0..3: W_N^(k + j Ms) = W_N^k W_8^j; then the inverse radix-4 pre-pass and W_M^(-k r)
        const T h = (T)0.70710678118654752440;
#pragma unroll
        for (int i = 0; i < E; i++) {
            const int k = t + i * NT;
            const C wk = tw[k << tw_shift_n];                                  // the bins are N/(8E) apart: one look-up each
            const C w1 = mk<T>((wk.x + wk.y) * h, (wk.y - wk.x) * h);          // W_N^k (1 - i)/sqrt2
            const C w2 = mk<T>(wk.y, -wk.x);                                   // -i W_N^k
            const C w3 = mk<T>((wk.y - wk.x) * h, (-wk.x - wk.y) * h);         // W_N^k (-1 - i)/sqrt2
            C z0, z1, z2, z3;
            if (HEAD) {
                z0 = inv_elem<T>(in, layout, k, M, sc, wk, hx, hh); z1 = inv_elem<T>(in, layout, k + MS, M, sc, w1, hx, hh);
                z2 = inv_elem<T>(in, layout, k + 2 * MS, M, sc, w2, hx, hh); z3 = inv_elem<T>(in, layout, k + 3 * MS, M, sc, w3, hx, hh);
            } else {
                z0 = inv_elem<T>(in, layout, k, M, sc, wk); z1 = inv_elem<T>(in, layout, k + MS, M, sc, w1);
                z2 = inv_elem<T>(in, layout, k + 2 * MS, M, sc, w2); z3 = inv_elem<T>(in, layout, k + 3 * MS, M, sc, w3);
            }
            const C y = dif4_residue<true>(r, z0, z1, z2, z3);
            v[i] = r == 0 ? y : cmul(y, cconj(cpow_small(cmul(wk, wk), r)));  // W_M^k = W_N^(2k)
        }
    } else {
#pragma unroll
    for (int i = 0; i < E; i++) {
        const int k = t + i * NT;
        const C wk = cmul(wbase, thread_root<T, (R0 == 1 ? 2 : 4) * E>(i));   // W_N^k
        const C lo = HEAD ? inv_elem<T>(in, layout, k, M, sc, wk, hx, hh) : inv_elem<T>(in, layout, k, M, sc, wk);
        if (R0 == 1) {
            v[i] = lo;
        } else {
            const C wq = mk<T>(wk.y, -wk.x);                                              // W_N^(k + N/4) = -i W_N^k
            const C hi = HEAD ? inv_elem<T>(in, layout, k + MS, M, sc, wq, hx, hh) : inv_elem<T>(in, layout, k + MS, M, sc, wq);
            if (r == 0) v[i] = cadd(lo, hi);
            else v[i] = cmul(csub(lo, hi), cconj(cmul(wpre, thread_root<T, 2 * E>(i))));   // W_M^(-k)
        }
    }
    }
}

template <class T, int LOG2MS, int R0, int LOG2E = 4>
BFIR_HD void inv_load(int t, int bx, int r, cpx<T> (&v)[1 << LOG2E], const cpx<T> *__restrict__ tw, int tw_shift_n, const InvArgs &a)
{
    constexpr int M = (1 << LOG2MS) * R0;
    const T *in = (const T *)inv_in(a) + bx * a.in_stride_x;
    const T sc = (T)a.scale_in;
    const int cset = (a.head_x != NULL && a.head_map != NULL) ? a.head_map[bx] : bx;
    if (a.head_x != NULL && a.head_blocks[cset] > 0) {   // uniform per CTA
        const T *hx = (const T *)a.head_x + bx * a.head_x_stride + (long long)a.state->cur_slot * (2 * M);
        const T *hh = (const T *)a.head_h + cset * a.head_h_stride;
        inv_load_impl<T, LOG2MS, R0, true, LOG2E>(t, r, v, tw, tw_shift_n, in, a.in_layout, sc, hx, hh);
    } else {
        inv_load_impl<T, LOG2MS, R0, false, LOG2E>(t, r, v, tw, tw_shift_n, in, a.in_layout, sc, NULL, NULL);
    }
}

// engine output (OUT_RAW) for one thread, sample format known at compile time (see BFIR_FMT_SWITCH)
template <class T, int LOG2MS, int R0, int LOG2E, int FMT>
BFIR_HD void inv_store_raw(int t, int r, const cpx<T> (&v)[1 << LOG2E], uint8_t *raw, long long step, T ovf_max, OverflowAcc &acc)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E;
    T lg = (T)0;
    unsigned int novf = 0;
    if constexpr (FMT >= FMT_FLOAT_LE) {
#pragma unroll
        for (int i = 0; i < E / 2; i++) {
            uint8_t *p = raw + (long long)(2 * (R0 * (t + i * NT) + r)) * step;
            float_stats_nb<T>(v[i].x, ovf_max, lg, novf);
            float_stats_nb<T>(v[i].y, ovf_max, lg, novf);
            store_raw_real<T, FMT>(p, v[i].x);
            store_raw_real<T, FMT>(p + step, v[i].y);
        }
    } else {
        int32_t imin, imax;
        int_limits(FMT, imin, imax);
        const T rmin = (T)imin, rmax = (T)imax;
        int32_t il = 0;
#pragma unroll
        for (int i = 0; i < E / 2; i++) {
            uint8_t *p = raw + (long long)(2 * (R0 * (t + i * NT) + r)) * step;
            store_raw_int(p, FMT, quantise_nb<T>(half_up<T>(v[i].x), rmin, rmax, imin, imax, lg, novf, il));
            store_raw_int(p + step, FMT, quantise_nb<T>(half_up<T>(v[i].y), rmin, rmax, imin, imax, lg, novf, il));
        }
        if (il > acc.intlargest) acc.intlargest = il;
    }
    acc.n_overflows += novf;
    if ((double)lg > acc.largest) acc.largest = (double)lg;
}

// inverse, phase 1: v[i] = z[R0 n + r], n = t + i*NTs;  x[2m] = Re z[m], x[2m+1] = Im z[m]
template <class T, int LOG2MS, int R0, int LOG2E = 4>
BFIR_HD void inv_store(int t, int bx, int r, const cpx<T> (&v)[1 << LOG2E], const InvArgs &a, OverflowAcc &acc)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E, M = MS * R0, L = M;
    typedef cpx<T> C;
    if (a.out_mode == OUT_TIME) {
        C *out = (C *)((T *)inv_out(a) + bx * a.out_stride_x);
#pragma unroll
        for (int i = 0; i < E; i++) out[R0 * (t + i * NT) + r] = v[i];
        return;
    }
    if (t == 0 && r == 0 && a.state != NULL) {         // brutefir.cpp:316-321
        const T y0 = v[0].x;
        if (!(y0 - y0 == (T)0)) {
#ifdef __CUDA_ARCH__
            atomicMin(&a.state->first_bad_channel, bx);
            if (a.host_flag != NULL) *(volatile int *)a.host_flag = 1;
#else
            if (bx < a.state->first_bad_channel) a.state->first_bad_channel = bx;
#endif
        }
    }
    // only the first L samples are consumed (fftw_convolver.cpp:425-430): z index < M/2 <=> i < E/2
    if (a.out_mode == OUT_REAL_L) {
        C *out = (C *)((T *)inv_out(a) + (long long)bx * L);
#pragma unroll
        for (int i = 0; i < E / 2; i++) out[R0 * (t + i * NT) + r] = v[i];
        return;
    }
    // OUT_RAW
    const int rb = bx - a.raw_ch_base;
    const int stream = rb / a.ch_per_stream, ch = rb - stream * a.ch_per_stream;
    const int bytes = fmt_bytes(a.fmt);
    uint8_t *raw = (uint8_t *)inv_out(a) + (long long)stream * a.out_stride_x + (long long)ch * bytes;
    const long long step = (long long)a.ch_per_stream * bytes;
    BFIR_FMT_SWITCH(a.fmt, (inv_store_raw<T, LOG2MS, R0, LOG2E, FMT>(t, r, v, raw, step, (T)a.ovf_max, acc)))
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// De-interleaving through a thread-block cluster and distributed shared memory. One CTA owns one channel, but the
// raw blocks are interleaved [frame][channel]: a CTA that reads (writes) its own samples touches `bytes` of every
// C * bytes frame -- for 7.1 double 8 of every 64 bytes, 32 sectors per warp instruction at 25 % efficiency, and the
// load / store unit spends 28 % (forward) / 39 % (inverse) of the kernel on them
// (profiles/r02n_fft_f64_13_stall_profile.txt). Here the CTAs of Cc <= 4 neighbouring channels of one stream form a
// cluster (launch attribute), and every CTA moves 1/Cc of the FRAMES of all Cc channels with 16-byte global accesses
// that tile memory exactly:
//   forward: CTA r loads frames [r F, (r+1) F), F = L/Cc, of the Cc channels, decodes them and scatters them into the
//            channels' PLANAR previous-block rows; cluster barrier; every CTA reads its own row back (fwd_load_cluster).
//   inverse: every CTA stores its encoded samples planar into its FFT buffer (free after the last pass); cluster
//            barrier; CTA r gathers frames [r F, (r+1) F) of the Cc channels through distributed shared memory into
//            16-byte chunks and writes them to the interleaved block; cluster barrier before anybody exits.
// Both exchanges were measured both ways (cfg1 x 16, per launch): forward through L2 20.1 us / through DSMEM 23.2 /
// strided 22.7; inverse through L2 23.5 / through DSMEM 21.8 / strided 23.8 -- each side keeps its faster variant.
// A 16-byte chunk holds S = 16 / bytes consecutive samples of the (frame, channel)-ordered tile; launch_rfft_* only
// enables the path when a chunk never straddles a frame (S <= Cc) or the tile is the whole frame (Cc == C).
template <int B> struct raw_word;
template <> struct raw_word<2> { typedef unsigned short type; };
template <> struct raw_word<4> { typedef unsigned int type; };
template <> struct raw_word<8> { typedef unsigned long long type; };

__device__ __forceinline__ int cluster_log2(int c) { return c >= 4 ? 2 : 1; }
#define BFIR_MAX_CLUSTER 4

// the same shared-memory address in every CTA of the cluster
template <class P> __device__ __forceinline__ void cluster_peers(P *own, int Cc, P *(&peer)[BFIR_MAX_CLUSTER])
{
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
#pragma unroll
    for (int k = 0; k < BFIR_MAX_CLUSTER; k++) peer[k] = k < Cc ? cl.map_shared_rank(own, k) : own;
}
template <class P> __device__ __forceinline__ P *cluster_pick(P *const (&peer)[BFIR_MAX_CLUSTER], int k)
{
    P *p = peer[0];
    if (k == 1) p = peer[1];
    if (k == 2) p = peer[2];
    if (k == 3) p = peer[3];
    return p;
}

// inverse side, after the planar rows are complete: gather one slab of frames through DSMEM and write it interleaved
template <int B>
__device__ __forceinline__ void cluster_gather_store(int t, int nt, const uint8_t *const (&rows)[BFIR_MAX_CLUSTER], uint8_t *out_stream,
                                                     int C, int Cc, int cg0, int f0, int nchunks)
{
    constexpr int S = 16 / B;
    typedef typename raw_word<B>::type W;
    const int lc = cluster_log2(Cc);
    for (int q = t; q < nchunks; q += nt) {
        const int s = q * S;
        union { uint4 u; W w[S]; } pack;
#pragma unroll
        for (int j = 0; j < S; j++) {
            const int idx = s + j, fl = idx >> lc, c = idx & (Cc - 1);
            pack.w[j] = ((const W *)cluster_pick(rows, c))[f0 + fl];
        }
        uint8_t *dst = out_stream + ((long long)(f0 + (s >> lc)) * C + cg0 + (s & (Cc - 1))) * B;
        *(uint4 *)dst = pack.u;
    }
}

// forward side: one slab of frames -> decoded samples scattered into the channels' planar rows (global memory)
template <class T, int FMT>
__device__ __forceinline__ void cluster_scatter_load(int t, int nt, const uint8_t *in_stream, T *rows, long long row_stride,
                                                     int C, int Cc, int cg0, int f0, int nchunks)
{
    constexpr int B = FMT == FMT_S8 ? 1 : (FMT <= FMT_S16_BE ? 2 : (FMT <= FMT_S24_BE ? 3 : (FMT <= FMT_FLOAT_BE ? 4 : 8)));
    if constexpr (B == 2 || B == 4 || B == 8) {
        constexpr int S = 16 / B;
        const int lc = cluster_log2(Cc);
        for (int q = t; q < nchunks; q += nt) {
            const int s = q * S;
            union { uint4 u; uint8_t b[16]; } pack;
            pack.u = __ldg((const uint4 *)(in_stream + ((long long)(f0 + (s >> lc)) * C + cg0 + (s & (Cc - 1))) * B));
#pragma unroll
            for (int j = 0; j < S; j++) {
                const int idx = s + j, fl = idx >> lc, c = idx & (Cc - 1);
                rows[(long long)c * row_stride + (f0 + fl)] = load_raw<T>(pack.b + j * B, FMT);
            }
        }
    }
}

// forward load phase in cluster mode (R0 = 1, IN_RAW_PREV). The exchange goes through the channels' planar
// previous-block rows, i.e. L2 (measured against the DSMEM variant of the same phase: 20.1 against 23.2 us for
// cfg1 x 16; DSMEM delivers ~20 B/clk per SM, B300_MICROARCH.md): CTA r scatters its slab of frames into the Cc
// rows -- which the kernel has to write anyway for the next block --, cluster barrier (release / acquire at cluster
// scope), every CTA reads its own row back with L2 loads. `staged_prev` != NULL: the previous block is arriving in
// shared memory by bulk copy (barrier `bar`), else it is read from its planar row.
template <class T, int LOG2MS, int LOG2E>
__device__ __forceinline__ void fwd_load_cluster(int t, int bx, int by, cpx<T> (&v)[1 << LOG2E], const FwdArgs &a,
                                                 const cpx<T> *staged_prev, uint64_t *bar)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E, NH = E / 2, L = MS;
    typedef cpx<T> C2;
    namespace cg = cooperative_groups;
    const int Cc = a.cluster, C = a.ch_per_stream;
    const int stream = bx / C, ch = bx - stream * C;
    const int r = ch & (Cc - 1), cg0 = ch - r;           // r = rank in the cluster (launch_rfft_forward aligns the grid)
    const int bytes = fmt_bytes(a.fmt);
    const unsigned int par = a.prev_parity & 1u;
    T *rows = (T *)a.prev + ((long long)(par ^ 1u) * a.n_channels + (bx - r)) * L;      // planar rows of the cluster's channels (written)
    const uint8_t *in_stream = (const uint8_t *)a.in + (long long)stream * a.in_stride_x;
    const int F = L >> cluster_log2(Cc);
    BFIR_FMT_SWITCH(a.fmt, (cluster_scatter_load<T, FMT>(t, NT, in_stream, rows, (long long)L, C, Cc, cg0, r * F, (L * bytes) >> 4)))
    cg::this_cluster().sync();                          // every channel's current block is in its planar row
    const C2 *cur = (const C2 *)(rows + (long long)r * L);
    C2 hi[NH];
#pragma unroll
    for (int j = 0; j < NH; j++) {
        if constexpr (sizeof(T) == 8) { const double2 d = __ldcg((const double2 *)cur + (t + j * NT)); hi[j].x = d.x; hi[j].y = d.y; }
        else { const float2 f = __ldcg((const float2 *)cur + (t + j * NT)); hi[j].x = f.x; hi[j].y = f.y; }
    }
    if (staged_prev != NULL) {
        mbar_wait(bar, 0);
#pragma unroll
        for (int j = 0; j < NH; j++) v[j] = staged_prev[t + j * NT];
    } else {
        const C2 *prev_rd = (const C2 *)((const T *)a.prev + ((long long)par * a.n_channels + bx) * L);
#pragma unroll
        for (int j = 0; j < NH; j++) v[j] = prev_rd[t + j * NT];
    }
#pragma unroll
    for (int j = 0; j < NH; j++) v[j + NH] = hi[j];
    fwd_bookkeeping<int>(t, bx, by, 0, a);
}

// inverse store phase in cluster mode (R0 = 1, OUT_RAW); smem_raw: the CTA's FFT buffer, free after the last pass
template <class T, int LOG2MS, int LOG2E>
__device__ __forceinline__ void inv_store_cluster(int t, int bx, const cpx<T> (&v)[1 << LOG2E], const InvArgs &a, OverflowAcc &acc,
                                                  unsigned char *smem_raw)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, NT = MS / E, L = MS;
    namespace cg = cooperative_groups;
    if (t == 0 && a.state != NULL) {                    // brutefir.cpp:316-321
        const T y0 = v[0].x;
        if (!(y0 - y0 == (T)0)) {
            atomicMin(&a.state->first_bad_channel, bx);
            if (a.host_flag != NULL) *(volatile int *)a.host_flag = 1;
        }
    }
    const int bytes = fmt_bytes(a.fmt);
    __syncthreads();                                    // every thread is past its last read of the FFT buffer
    BFIR_FMT_SWITCH(a.fmt, (inv_store_raw<T, LOG2MS, 1, LOG2E, FMT>(t, 0, v, (uint8_t *)smem_raw, (long long)bytes, (T)a.ovf_max, acc)))
    cg::this_cluster().sync();                          // the planar rows of all Cc channels are complete
    const int Cc = a.cluster, C = a.ch_per_stream;
    const int rb = bx - a.raw_ch_base;
    const int stream = rb / C, ch = rb - stream * C;
    const int r = ch & (Cc - 1), cg0 = ch - r;
    const int F = L >> cluster_log2(Cc);
    uint8_t *out_stream = (uint8_t *)inv_out(a) + (long long)stream * a.out_stride_x;
    const uint8_t *rows[BFIR_MAX_CLUSTER];
    cluster_peers<const uint8_t>((const uint8_t *)smem_raw, Cc, rows);
    const int nchunks = (L * bytes) >> 4;
    if (bytes == 8) cluster_gather_store<8>(t, NT, rows, out_stream, C, Cc, cg0, r * F, nchunks);
    else if (bytes == 4) cluster_gather_store<4>(t, NT, rows, out_stream, C, Cc, cg0, r * F, nchunks);
    else cluster_gather_store<2>(t, NT, rows, out_stream, C, Cc, cg0, r * F, nchunks);
    cg::this_cluster().sync();                          // nobody leaves while a peer still reads its rows
}

// merge per-thread statistics into the channel's counters: max / sum are order independent, so the
// result equals the reference's serial update (real2raw.cpp:17-32, dither.cpp:226-271)
__device__ __forceinline__ void overflow_commit(OverflowStats *dst, OverflowAcc acc)
{
    unsigned long long lb = (unsigned long long)__double_as_longlong(acc.largest);
    if (blockDim.x >= 32) { // whole warps only (blocks below 32 threads exist for L < 512)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.n_overflows += __shfl_xor_sync(0xffffffffu, acc.n_overflows, o);
            acc.intlargest = max(acc.intlargest, __shfl_xor_sync(0xffffffffu, acc.intlargest, o));
            unsigned long long other = __shfl_xor_sync(0xffffffffu, lb, o);
            lb = lb > other ? lb : other;
        }
        if ((threadIdx.x & 31) != 0) return;
    }
    if (acc.n_overflows) atomicAdd(&dst->n_overflows, acc.n_overflows);
    if (acc.intlargest > 0) atomicMax(&dst->intlargest, acc.intlargest);
    if (lb) atomicMax(&dst->largest_bits, lb);
}

// 8 points per thread at 4096 points per CTA: two CTAs of 512 threads per SM (64 registers each)
#ifndef BFIR_E8_MINB
#define BFIR_E8_MINB(log2e, log2ms) (((log2e) == 3 && (log2ms) == 12) ? 2 : 0)
#endif
// grid = (buffers, partitions, R0); tw_shift_m = log2(table length / Ms), tw_shift_n = log2(table length / N)
// CL: the cluster variant (de-interleaving through a thread-block cluster, R0 = 1 only) is its own instantiation, so that
// the plain kernels keep their register allocation
template <class T, int LOG2MS, int R0, int LOG2E = 4, bool CL = false>
__global__ void __launch_bounds__((1 << LOG2MS) >> LOG2E, BFIR_E8_MINB(LOG2E, LOG2MS)) rfft_forward_kernel(const FwdArgs a, const cpx<T> *__restrict__ tw, int tw_shift_m, int tw_shift_n)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx<T> *smem = reinterpret_cast<cpx<T> *>(smem_raw);
    const int t = threadIdx.x, bx = blockIdx.x + a.ch_base, by = blockIdx.y, r = blockIdx.z;
    cpx<T> v[1 << LOG2E];
    bool staged = false;
    if constexpr (R0 == 1) staged = a.tma != 0;
    constexpr bool clustered = CL && R0 == 1;
    if constexpr (R0 == 1) if (staged) {
        // previous block: L reals, contiguous in the planar ping-pong buffer -> head of the (still idle) FFT buffer
        __shared__ __align__(8) uint64_t bar;
        constexpr int L = 1 << LOG2MS;
        if (t == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (t == 0) bulk_stage(smem_raw, (const T *)a.prev + ((long long)(a.prev_parity & 1u) * a.n_channels + bx) * L, (uint32_t)(L * sizeof(T)), &bar);
        if constexpr (clustered) fwd_load_cluster<T, LOG2MS, LOG2E>(t, bx, by, v, a, smem, &bar);
        else fwd_load<T, LOG2MS, R0, LOG2E, true>(t, bx, by, r, v, tw, tw_shift_n, a, smem, &bar);
        __syncthreads();                               // every thread has read its part of the staged block
    }
    if constexpr (clustered) { if (!staged) fwd_load_cluster<T, LOG2MS, LOG2E>(t, bx, by, v, a, NULL, NULL); }
    else { if (!staged) fwd_load<T, LOG2MS, R0, LOG2E>(t, bx, by, r, v, tw, tw_shift_n, a); }
    fft_passes<T, LOG2MS, false, 0, 0, LOG2E>::run(t, v, smem, tw, tw_shift_m);
    BlockFFT<T, LOG2MS, false, LOG2E>::store_natural(t, v, smem);
    __syncthreads();
    if constexpr (R0 == 4) {
        // launched as clusters of four CTAs along z (rank = r): residues 1 and 3 read each other's sub-transform
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();                                 // all four sub-transforms are in shared memory
        const cpx<T> *partner = (r == 1 || r == 3) ? cluster.map_shared_rank(smem, 4 - r) : smem;
        fwd_split_store<T, LOG2MS, R0, LOG2E>(t, bx, by, r, smem, tw, tw_shift_n, a, partner);
        cluster.sync();                                 // nobody leaves while a partner still reads its buffer
    } else {
        fwd_split_store<T, LOG2MS, R0, LOG2E>(t, bx, by, r, smem, tw, tw_shift_n, a);
    }
}

template <class T, int LOG2MS, int R0, int LOG2E = 4, bool CL = false>
__global__ void __launch_bounds__((1 << LOG2MS) >> LOG2E, BFIR_E8_MINB(LOG2E, LOG2MS)) rfft_inverse_kernel(const InvArgs a, const cpx<T> *__restrict__ tw, int tw_shift_m, int tw_shift_n)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx<T> *smem = reinterpret_cast<cpx<T> *>(smem_raw);
    const int t = threadIdx.x, bx = blockIdx.x + a.ch_base, r = blockIdx.z;
    cpx<T> v[1 << LOG2E];
    bool staged = false;
    if constexpr (R0 == 1) staged = a.tma != 0;
    if constexpr (R0 == 1) if (staged) {
        // input spectrum: N reals of this channel, contiguous -> the (still idle) FFT buffer, read back bin by bin
        __shared__ __align__(8) uint64_t bar;
        constexpr int N = 2 << LOG2MS;
        if (t == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (t == 0) bulk_stage(smem_raw, (const T *)inv_in(a) + bx * a.in_stride_x, (uint32_t)(N * sizeof(T)), &bar);
        mbar_wait(&bar, 0);
        inv_load_impl<T, LOG2MS, R0, false, LOG2E>(t, r, v, tw, tw_shift_n, reinterpret_cast<const T *>(smem_raw), a.in_layout, (T)a.scale_in, NULL, NULL);
        __syncthreads();                               // every thread has read its bins before the passes overwrite them
    }
    if (!staged) inv_load<T, LOG2MS, R0, LOG2E>(t, bx, r, v, tw, tw_shift_n, a);
    fft_passes<T, LOG2MS, true, 0, 0, LOG2E>::run(t, v, smem, tw, tw_shift_m);
    OverflowAcc acc;
    acc.n_overflows = 0; acc.intlargest = 0; acc.largest = 0.0;
    if constexpr (CL && R0 == 1) inv_store_cluster<T, LOG2MS, LOG2E>(t, bx, v, a, acc, smem_raw);
    else inv_store<T, LOG2MS, R0, LOG2E>(t, bx, r, v, a, acc);
    if (a.out_mode == OUT_RAW && a.stats != NULL) overflow_commit(&a.stats[bx], acc);
    if (a.state != NULL && blockIdx.x == 0 && blockIdx.y == 0 && t == 0 && r == 0) a.state->blockcounter += a.n_multi > 0 ? (unsigned int)a.n_multi : 1u; // brutefir.cpp:337-340
}
#endif

} // namespace bfir
