// Frequency-domain multiply-accumulate over the filter partitions -- THE bandwidth-bound kernel.
//
// Reference: brutefir::run's partition loop, brutefir.cpp:288-299, i.e. one convolver_convolve
// (fftw_convolver.cpp:1465-1493 / 2161-2189) followed by P-1 convolver_convolve_add
// (:1497-1525 / :2192-2220), each of which re-reads and re-writes the accumulator. Here the
// accumulator lives in registers for the whole partition loop, so one channel-block moves exactly
//     B_mac = (2 * P_eff + 1) * N * realsize   bytes
// (P_eff coefficient spectra + P_eff delay-line spectra in, one accumulated spectrum out).
//
// All spectra are in the reference's ORD layout (groups of 8 reals = [Re k..k+3 | Im k..k+3]); the
// real DC and Nyquist bins share group 0 (slots 0 and 4) and are accumulated separately exactly like
// the reference's d1s/d2s (fftw_convolver.cpp:1475-1476, 1507-1508).
#pragma once
#include "rfft_kernels.cuh"

namespace bfir {

struct MacArgs {
    const void *fdl;         // [channels][n_slots][N]   frequency-domain delay line, slot = blockcounter % n_slots
    const void *coeffs;      // [channels][coeff_alloc][N]
    void *acc;               // [channels][N]
    long long fdl_stride_ch, coeff_stride_ch; // elements
    int N;
    int n_slots;             // delay-line slots per channel (filter_blocks + 1 in the engine)
    int n_parts;             // filter_blocks
    int part_begin, part_count; // partition shard convolved by this launch (whole filter: 0, n_slots)
    const int *coeff_blocks; // [channels] coefficient partitions actually loaded
    const int *coeff_map;    // [channels] coefficient set each filter channel convolves with; NULL = its own (bfir_set_coeff_map)
    const int *procblocks;   // [channels] blocks seen so far, already counting the current one
    const EngineState *state;
    int block_offset;        // 0: blockcounter is the current block; used by tests
    int ch_base;             // first channel of this launch (channel-group pipelining)
    void *acc_next;          // pair kernel: accumulated spectrum of block t+1, [channels][N]
    void *acc_multi[8];      // multi / wide kernel: accumulated spectra of blocks t .. t+NB-1
    int use_abs_block;       // 1: block index t = abs_block, given by the host (stage pipeline) instead of the device counter
    unsigned int abs_block;
    int procblocks_bias;     // 1: look-ahead launch for the NEXT block, whose forward transform has not counted itself yet
    PeerPush push;           // enabled: partial sums go to the owner rank's receive buffer (fused reduce)
    int push_phase;          // multi kernel: receive-buffer phase of block t (block t+b: push_phase + b)
    int tiles_x, n_ch_launch; // persistent eight-block kernel: tiles per channel, channels of the launch
};

template <class T> struct vec8 { T v[8]; };

#ifdef __CUDACC__
__device__ __forceinline__ void ld8(const float *p, float (&r)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void ld8(const double *p, double (&r)[8])
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y; r[4] = c.x; r[5] = c.y; r[6] = d.x; r[7] = d.y;
}
__device__ __forceinline__ void st8(float *p, const float (&r)[8])
{
    reinterpret_cast<float4 *>(p)[0] = make_float4(r[0], r[1], r[2], r[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(r[4], r[5], r[6], r[7]);
}
__device__ __forceinline__ void st8(double *p, const double (&r)[8])
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(r[0], r[1]); q[1] = make_double2(r[2], r[3]);
    q[2] = make_double2(r[4], r[5]); q[3] = make_double2(r[6], r[7]);
}

template <class T>
__device__ __forceinline__ void mac8(T (&acc)[8], const T (&b)[8], const T (&c)[8])
{
#pragma unroll
    for (int j = 0; j < 4; j++) {
        acc[j] = fma(b[j], c[j], acc[j]);
        acc[j] = fma(-b[j + 4], c[j + 4], acc[j]);
        acc[j + 4] = fma(b[j], c[j + 4], acc[j + 4]);
        acc[j + 4] = fma(b[j + 4], c[j], acc[j + 4]);
    }
}

// grid: (ceil(N/8 / GPC), channels), 256 threads. The CTA covers GPC = 256/SPLIT consecutive ORD groups
// (4 complex bins each); its threads form SPLIT slices, slice s accumulating partitions i = s (mod SPLIT)
// of its group in registers, UNROLL partitions in flight at a time. The slices are then summed through
// shared memory in slice order (deterministic). SPLIT trades per-thread work for CTA count: SPLIT 1 is
// one thread per group over all partitions; larger SPLIT gives few-channel configurations enough CTAs
// to fill 148 SMs and shortens the tail wave of large ones.
template <class T, int SPLIT, int UNROLL>
__global__ void __launch_bounds__(256) partition_mac_kernel(const MacArgs a)
{
    constexpr int GPC = 256 / SPLIT;
    const int slice = threadIdx.x / GPC, gl = threadIdx.x - slice * GPC;
    const int g = blockIdx.x * GPC + gl;
    const int ch = blockIdx.y + a.ch_base;
    const bool active = g * 8 < a.N;
    const unsigned int t = a.state->blockcounter + (unsigned int)a.block_offset;
    const int cs_ = a.coeff_map ? a.coeff_map[ch] : ch;
    const int peff = min(a.coeff_blocks[cs_], min(a.procblocks[ch] + a.procblocks_bias, a.n_parts)); // brutefir.cpp:292, 265-268
    const int i_end = min(peff, a.part_begin + a.part_count);
    const T *fdl = (const T *)a.fdl + ch * a.fdl_stride_ch + (long long)g * 8;
    const T *cf = (const T *)a.coeffs + cs_ * a.coeff_stride_ch + (long long)g * 8;
    const unsigned int P = (unsigned int)a.n_slots;

    T acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = (T)0;
    T dc = (T)0, ny = (T)0;

    if (active) {
        int i = a.part_begin + slice;
        for (; i + (UNROLL - 1) * SPLIT < i_end; i += UNROLL * SPLIT) {
            T b[UNROLL][8], c[UNROLL][8];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const unsigned int slot = (t - (unsigned int)(i + u * SPLIT)) % P; // brutefir.cpp:294
                ld8(fdl + (long long)slot * a.N, b[u]);
                ld8(cf + (long long)(i + u * SPLIT) * a.N, c[u]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (g == 0) { dc = fma(b[u][0], c[u][0], dc); ny = fma(b[u][4], c[u][4], ny); }
                mac8<T>(acc, b[u], c[u]);
            }
        }
        for (; i < i_end; i += SPLIT) {
            T b[8], c[8];
            const unsigned int slot = (t - (unsigned int)i) % P;
            ld8(fdl + (long long)slot * a.N, b);
            ld8(cf + (long long)i * a.N, c);
            if (g == 0) { dc = fma(b[0], c[0], dc); ny = fma(b[4], c[4], ny); }
            mac8<T>(acc, b, c);
        }
        if (g == 0) { acc[0] = dc; acc[4] = ny; }
    }
    if (SPLIT > 1) {
        __shared__ T red[SPLIT > 1 ? (SPLIT - 1) * GPC * 8 : 1];
        if (slice > 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) red[((slice - 1) * 8 + j) * GPC + gl] = acc[j];
        }
        __syncthreads();
        if (slice > 0) return;
#pragma unroll
        for (int s = 1; s < SPLIT; s++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] += red[((s - 1) * 8 + j) * GPC + gl];
    }
    if (active) {
        T *dst = a.push.enabled ? peer_dst<T>(a.push, ch, a.N, t & 1u) : (T *)a.acc + (long long)ch * a.N;
        st8(dst + (long long)g * 8, acc);
    }
}

// Two consecutive blocks t, t+1 in one pass (offline / pipelined callers that have the next block at hand; steady
// state only: every channel has seen at least n_slots blocks, so P_eff = coeff_blocks). H[i] is read once for both
// blocks, and X[t-i] serves block t at partition i and block t+1 at partition i+1, so a contiguous run of partitions
// costs ONE coefficient and ONE delay-line load per partition: per channel (2 P_eff + SPLIT + 2) N rs bytes for two
// blocks instead of 2 (2 P_eff + 1) N rs. Both forward transforms have run (block t+1 sits in slot (t+1) % P).
// Slice s owns the contiguous partitions [s cs, (s+1) cs), cs = ceil(P_eff / SPLIT); slices are summed in order.
template <class T, int SPLIT, int UNROLL>
__global__ void __launch_bounds__(256) partition_mac_pair_kernel(const MacArgs a)
{
    constexpr int GPC = 256 / SPLIT;
    const int slice = threadIdx.x / GPC, gl = threadIdx.x - slice * GPC;
    const int g = blockIdx.x * GPC + gl;
    const int ch = blockIdx.y + a.ch_base;
    const bool active = g * 8 < a.N;
    const unsigned int t = a.use_abs_block ? a.abs_block : a.state->blockcounter + (unsigned int)a.block_offset;
    const int cs_ = a.coeff_map ? a.coeff_map[ch] : ch;
    const int peff = min(a.coeff_blocks[cs_], a.n_parts);
    const int cs = (peff + SPLIT - 1) / SPLIT;
    const int i0 = slice * cs, i1 = min(peff, i0 + cs);
    const T *fdl = (const T *)a.fdl + ch * a.fdl_stride_ch + (long long)g * 8;
    const T *cf = (const T *)a.coeffs + cs_ * a.coeff_stride_ch + (long long)g * 8;
    const unsigned int P = (unsigned int)a.n_slots;

    T acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { acc0[j] = (T)0; acc1[j] = (T)0; }
    T dc0 = (T)0, ny0 = (T)0, dc1 = (T)0, ny1 = (T)0;

    if (active && i0 < i1) {
        T xa[8];                                                   // X[t+1-i]: block t+1's operand at partition i
        ld8(fdl + (long long)((t + 1u - (unsigned int)i0) % P) * a.N, xa);
        int i = i0;
        for (; i + UNROLL <= i1; i += UNROLL) {
            T xb[UNROLL][8], c[UNROLL][8];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                ld8(fdl + (long long)((t - (unsigned int)(i + u)) % P) * a.N, xb[u]);   // X[t-i]: block t at i, block t+1 at i+1
                ld8(cf + (long long)(i + u) * a.N, c[u]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (g == 0) {
                    dc1 = fma(xa[0], c[u][0], dc1); ny1 = fma(xa[4], c[u][4], ny1);
                    dc0 = fma(xb[u][0], c[u][0], dc0); ny0 = fma(xb[u][4], c[u][4], ny0);
                }
                mac8<T>(acc1, xa, c[u]);
                mac8<T>(acc0, xb[u], c[u]);
#pragma unroll
                for (int j = 0; j < 8; j++) xa[j] = xb[u][j];
            }
        }
        for (; i < i1; i++) {
            T xb[8], c[8];
            ld8(fdl + (long long)((t - (unsigned int)i) % P) * a.N, xb);
            ld8(cf + (long long)i * a.N, c);
            if (g == 0) {
                dc1 = fma(xa[0], c[0], dc1); ny1 = fma(xa[4], c[4], ny1);
                dc0 = fma(xb[0], c[0], dc0); ny0 = fma(xb[4], c[4], ny0);
            }
            mac8<T>(acc1, xa, c);
            mac8<T>(acc0, xb, c);
#pragma unroll
            for (int j = 0; j < 8; j++) xa[j] = xb[j];
        }
        if (g == 0) { acc0[0] = dc0; acc0[4] = ny0; acc1[0] = dc1; acc1[4] = ny1; }
    }
    if (SPLIT > 1) {
        __shared__ T red[SPLIT > 1 ? (SPLIT - 1) * GPC * 16 : 1];
        if (slice > 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                red[((slice - 1) * 16 + j) * GPC + gl] = acc0[j];
                red[((slice - 1) * 16 + 8 + j) * GPC + gl] = acc1[j];
            }
        }
        __syncthreads();
        if (slice > 0) return;
#pragma unroll
        for (int s = 1; s < SPLIT; s++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc0[j] += red[((s - 1) * 16 + j) * GPC + gl];
                acc1[j] += red[((s - 1) * 16 + 8 + j) * GPC + gl];
            }
    }
    if (active) {
        st8((T *)a.acc + (long long)ch * a.N + (long long)g * 8, acc0);
        st8((T *)a.acc_next + (long long)ch * a.N + (long long)g * 8, acc1);
    }
}

// NB consecutive blocks t .. t+NB-1 in one pass (the pair kernel generalised; single precision has the registers
// for NB = 4): a window of NB delay-line spectra slides over a contiguous run of partitions, so every partition
// costs one coefficient and one delay-line load for NB multiply-accumulates: per channel
// (2 P_eff + (NB-1) SPLIT + NB) N rs bytes for NB blocks instead of NB (2 P_eff + 1) N rs.
// Block t+b uses X[t+b-i] at partition i; all NB forward transforms have run (needs P + NB - 1 delay-line slots).
template <class T, int NB, int SPLIT, int THREADS = 256>
__global__ void __launch_bounds__(THREADS) partition_mac_multi_kernel(const MacArgs a)
{
    constexpr int GPC = THREADS / SPLIT;
    const int slice = threadIdx.x / GPC, gl = threadIdx.x - slice * GPC;
    const int g = blockIdx.x * GPC + gl;
    const int ch = blockIdx.y + a.ch_base;
    const bool active = g * 8 < a.N;
    const unsigned int t = a.use_abs_block ? a.abs_block : a.state->blockcounter + (unsigned int)a.block_offset;
    const int cs_ = a.coeff_map ? a.coeff_map[ch] : ch;
    const int peff = min(a.coeff_blocks[cs_], a.n_parts);
    // partition shard [part_begin, part_begin + part_count) (the whole filter: 0, n_parts), dealt to the slices
    const int p_end = min(peff, a.part_begin + a.part_count);
    const int cs = (max(p_end - a.part_begin, 0) + SPLIT - 1) / SPLIT;
    const int i0 = a.part_begin + slice * cs, i1 = min(p_end, i0 + cs);
    const T *fdl = (const T *)a.fdl + ch * a.fdl_stride_ch + (long long)g * 8;
    const T *cf = (const T *)a.coeffs + cs_ * a.coeff_stride_ch + (long long)g * 8;
    const unsigned int P = (unsigned int)a.n_slots;

    T acc[NB][8];
    T dc[NB], ny[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        dc[b] = (T)0; ny[b] = (T)0;
#pragma unroll
        for (int j = 0; j < 8; j++) acc[b][j] = (T)0;
    }
    if (active && i0 < i1) {
        T x[NB][8];                                                // x[b] = X[t+b-i]: block t+b's operand at partition i
#pragma unroll
        for (int b = 1; b < NB; b++) ld8(fdl + (long long)((t + (unsigned int)b - (unsigned int)i0) % P) * a.N, x[b]);
#pragma unroll 2
        for (int i = i0; i < i1; i++) {
            T c[8];
            ld8(fdl + (long long)((t - (unsigned int)i) % P) * a.N, x[0]);
            ld8(cf + (long long)i * a.N, c);
#pragma unroll
            for (int b = 0; b < NB; b++) {
                if (g == 0) { dc[b] = fma(x[b][0], c[0], dc[b]); ny[b] = fma(x[b][4], c[4], ny[b]); }
                mac8<T>(acc[b], x[b], c);
            }
#pragma unroll
            for (int b = NB - 1; b > 0; b--)                       // next partition: block t+b needs what block t+b-1 just used
#pragma unroll
                for (int j = 0; j < 8; j++) x[b][j] = x[b - 1][j];
        }
        if (g == 0) {
#pragma unroll
            for (int b = 0; b < NB; b++) { acc[b][0] = dc[b]; acc[b][4] = ny[b]; }
        }
    }
    if (SPLIT > 1) {
        __shared__ T red[SPLIT > 1 ? (SPLIT - 1) * GPC * 8 * NB : 1];
        if (slice > 0) {
#pragma unroll
            for (int b = 0; b < NB; b++)
#pragma unroll
                for (int j = 0; j < 8; j++) red[((slice - 1) * 8 * NB + b * 8 + j) * GPC + gl] = acc[b][j];
        }
        __syncthreads();
        if (slice > 0) return;
#pragma unroll
        for (int s = 1; s < SPLIT; s++)
#pragma unroll
            for (int b = 0; b < NB; b++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc[b][j] += red[((s - 1) * 8 * NB + b * 8 + j) * GPC + gl];
    }
    if (active) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
            T *dst = a.push.enabled ? peer_dst<T>(a.push, ch, a.N, (unsigned int)(a.push_phase + b)) : (T *)a.acc_multi[b] + (long long)ch * a.N;
            st8(dst + (long long)g * 8, acc[b]);
        }
    }
}

// EIGHT consecutive blocks per launch (bfir_run_device_oct): per channel (2 P_eff + 15) N rs bytes for eight blocks, i.e.
// 9.9 spectra per block at P = 32 where four blocks per launch move 17.75 and one block per launch 65. Eight accumulators
// and an eight-deep window of delay-line spectra do not fit the registers of a thread that owns a whole ORD group in
// double precision, so a thread owns W reals of a group: W = 4 in single precision (bins 2h, 2h+1 of the group =
// [Re, Re | Im, Im]: two 8-byte loads per spectrum), W = 2 in double precision (one bin: two 8-byte loads; 148 registers,
// three CTAs of 128 threads per SM, the eight partitions of a loop body requested before the first is used). Wider
// threads were measured and are slower (fewer CTAs per SM, fewer loads in flight). The window is a circular buffer indexed at
// compile time: the body handles eight partitions, partition i+u loading X[t-i-u] into slot (-u) & 7 -- the slot whose
// spectrum X[t-i+8-u] no later step needs -- and block b reading slot (b-u) & 7, so nothing is ever moved.
// One slice per group (large batches); grid (ceil(N/W / THREADS), channels).
template <class T, int W> __device__ __forceinline__ void ldw(const T *p, T (&r)[W])
{
    if constexpr (W == 8) { ld8(p, r); }
    else if constexpr (W == 2) { r[0] = __ldg(p); r[1] = __ldg(p + 4); }
    else if constexpr (sizeof(T) == 8) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p + 4));
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
    } else {
        const float2 a = __ldg(reinterpret_cast<const float2 *>(p)), b = __ldg(reinterpret_cast<const float2 *>(p + 4));
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
    }
}
template <class T, int W> __device__ __forceinline__ void stw(T *p, const T (&r)[W])
{
    if constexpr (W == 8) { st8(p, r); }
    else if constexpr (W == 2) { p[0] = r[0]; p[4] = r[1]; }
    else if constexpr (sizeof(T) == 8) {
        *reinterpret_cast<double2 *>(p) = make_double2(r[0], r[1]); *reinterpret_cast<double2 *>(p + 4) = make_double2(r[2], r[3]);
    } else {
        *reinterpret_cast<float2 *>(p) = make_float2(r[0], r[1]); *reinterpret_cast<float2 *>(p + 4) = make_float2(r[2], r[3]);
    }
}
template <class T, int W> __device__ __forceinline__ void macw(T (&acc)[W], const T (&b)[W], const T (&c)[W])
{
    constexpr int H = W / 2;
#pragma unroll
    for (int j = 0; j < H; j++) {
        acc[j] = fma(b[j], c[j], acc[j]);
        acc[j] = fma(-b[j + H], c[j + H], acc[j]);
        acc[j + H] = fma(b[j], c[j + H], acc[j + H]);
        acc[j + H] = fma(b[j + H], c[j], acc[j + H]);
    }
}

template <class T, int W, int AHEAD>
__device__ __forceinline__ void mac_oct_tile(const MacArgs &a, int q, int ch)
{
    constexpr int NB = 8, HALVES = 8 / W, H = W / 2;
    const int g = q / HALVES, h = q - g * HALVES;
    if (g * 8 >= a.N) return;
    const unsigned int t = a.use_abs_block ? a.abs_block : a.state->blockcounter + (unsigned int)a.block_offset;
    const int cs_ = a.coeff_map ? a.coeff_map[ch] : ch;
    const int peff = min(a.coeff_blocks[cs_], a.n_parts);
    const int i0 = a.part_begin, i1 = min(peff, a.part_begin + a.part_count);
    const long long off = (long long)g * 8 + h * H;                 // W = 4: reals [2h, 2h+1] and [4+2h, 4+2h+1] of the group
    const T *fdl = (const T *)a.fdl + ch * a.fdl_stride_ch + off;
    const T *cf = (const T *)a.coeffs + cs_ * a.coeff_stride_ch + off;
    const unsigned int P = (unsigned int)a.n_slots;

    T acc[NB][W], xs[NB][W];
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
        for (int j = 0; j < W; j++) acc[b][j] = (T)0;
    if (i0 < i1) {
#pragma unroll
        for (int d = 1; d < NB; d++) ldw<T, W>(fdl + (long long)((t + (unsigned int)d - (unsigned int)i0) % P) * a.N, xs[d]);
        for (int i = i0; i < i1; i += NB) {
#pragma unroll
            for (int grp = 0; grp < NB / AHEAD; grp++) {
                // AHEAD partitions' operands are requested before the first of them is used: with two CTAs of 128 threads
                // per SM that is 64 KB in flight per SM, what the four-block kernel keeps in flight too
                T xn[AHEAD][W], cn[AHEAD][W];
#pragma unroll
                for (int k = 0; k < AHEAD; k++) {
                    const int u = grp * AHEAD + k;
                    if (i + u < i1) {
                        ldw<T, W>(fdl + (long long)((t - (unsigned int)(i + u)) % P) * a.N, xn[k]);
                        ldw<T, W>(cf + (long long)(i + u) * a.N, cn[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < AHEAD; k++) {
                    const int u = grp * AHEAD + k;
                    if (i + u < i1) {
#pragma unroll
                        for (int j = 0; j < W; j++) xs[(NB - u) & (NB - 1)][j] = xn[k][j];
#pragma unroll
                        for (int b = 0; b < NB; b++) macw<T, W>(acc[b], xs[(b - u + NB) & (NB - 1)], cn[k]);
                    }
                }
            }
        }
    }
    T *dst[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        dst[b] = a.push.enabled ? peer_dst<T>(a.push, ch, a.N, (unsigned int)(a.push_phase + b)) : (T *)a.acc_multi[b] + (long long)ch * a.N;
        stw<T, W>(dst[b] + off, acc[b]);
    }
    // The real DC and Nyquist bins (slots 0 and 4 of group 0) are products of reals (fftw_convolver.cpp:1507-1508), not
    // the complex product the thread above has just stored for "bin 0": the one thread that owns them walks the
    // partitions once more with scalars and overwrites the two slots (same accumulation order as the other kernels).
    if (g == 0 && h == 0 && i0 < i1) {
        T d0[NB], d4[NB], w0[NB], w4[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) { d0[b] = (T)0; d4[b] = (T)0; w0[b] = (T)0; w4[b] = (T)0; }
#pragma unroll
        for (int d = 1; d < NB; d++) {
            const T *x = fdl + (long long)((t + (unsigned int)d - (unsigned int)i0) % P) * a.N;
            w0[d] = __ldg(x); w4[d] = __ldg(x + 4);
        }
        for (int i = i0; i < i1; i++) {
            const T *x = fdl + (long long)((t - (unsigned int)i) % P) * a.N, *c = cf + (long long)i * a.N;
            w0[0] = __ldg(x); w4[0] = __ldg(x + 4);
            const T c0 = __ldg(c), c4 = __ldg(c + 4);
#pragma unroll
            for (int b = 0; b < NB; b++) { d0[b] = fma(w0[b], c0, d0[b]); d4[b] = fma(w4[b], c4, d4[b]); }
#pragma unroll
            for (int b = NB - 1; b > 0; b--) { w0[b] = w0[b - 1]; w4[b] = w4[b - 1]; }
        }
#pragma unroll
        for (int b = 0; b < NB; b++) { dst[b][0] = d0[b]; dst[b][4] = d4[b]; }
    }
}

template <class T, int W, int THREADS, int AHEAD>
__global__ void __launch_bounds__(THREADS) partition_mac_oct_kernel(const MacArgs a)
{
    mac_oct_tile<T, W, AHEAD>(a, blockIdx.x * THREADS + threadIdx.x, blockIdx.y + a.ch_base);
}

// The same as a PERSISTENT grid of a.persist_ctas CTAs, each alone on its SM (the launch asks for most of the shared
// memory), walking the (channel, tile) list: the partition sum is bound by bytes in flight, not by SM cycles, and ~64 KB in
// flight on each of a THIRD of the SMs already saturates HBM -- so it can leave the other SMs to the transform kernels of
// the neighbouring calls (whole-SM CTAs, which otherwise wait until a sum's CTAs have drained from an SM).
template <class T, int W, int THREADS, int AHEAD>
__global__ void __launch_bounds__(THREADS) partition_mac_oct_persistent_kernel(const MacArgs a)
{
    const int n_tiles = a.tiles_x * a.n_ch_launch;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int c = tile / a.tiles_x, xb = tile - c * a.tiles_x;
        mac_oct_tile<T, W, AHEAD>(a, xb * THREADS + threadIdx.x, c + a.ch_base);
    }
}

typedef void (*mac_kernel_t)(const MacArgs);
// reals per thread: BFIR_OCT_W overrides (measurement)
template <class T> inline int mac_oct_reals_per_thread()
{
    static const int forced = [] { const char *env = getenv("BFIR_OCT_W"); return env ? atoi(env) : 0; }();
    // measured (tools/kernel_times.py --octs): double cfg1 x 16 W = 2 0.229 ms per launch (0.88 of the HBM peak on the bytes
    // that must move), W = 4 0.326 (224 registers, two CTAs per SM); float cfg3 shape W = 4 0.575 ms (0.84), W = 8 0.766
    if (sizeof(T) == 8) return forced == 4 ? 4 : 2;
    return forced == 8 ? 8 : 4;
}
template <class T> inline mac_kernel_t mac_oct_persistent_kernel()
{
    const int w = mac_oct_reals_per_thread<T>();
    if constexpr (sizeof(T) == 8) { if (w == 2) return partition_mac_oct_persistent_kernel<T, 2, 256, 8>; return partition_mac_oct_persistent_kernel<T, 4, 256, 4>; }
    else { if (w == 4) return partition_mac_oct_persistent_kernel<T, 4, 256, 8>; return partition_mac_oct_persistent_kernel<T, 8, 256, 4>; }
}
template <class T> inline mac_kernel_t mac_oct_kernel()
{
    const int w = mac_oct_reals_per_thread<T>();
    if constexpr (sizeof(T) == 8) { if (w == 2) return partition_mac_oct_kernel<T, 2, 128, 8>; return partition_mac_oct_kernel<T, 4, 128, 4>; }
    else { if (w == 4) return partition_mac_oct_kernel<T, 4, 128, 8>; return partition_mac_oct_kernel<T, 8, 128, 4>; }
}
// four blocks per launch. Shared memory of the slice reduction: (SPLIT-1) * 256/SPLIT * 32 reals, i.e. <= 32 KB in
// single precision and 48 KB at SPLIT 4 in double precision (the static limit), so double stops at SPLIT 4
template <class T> inline mac_kernel_t mac_quad_kernel_for_split(int split, int threads = 256)
{
    if (sizeof(T) == 8) {
        // 128 threads: 214 registers x 128 = 27 K registers per CTA, so that a transform CTA (256 threads x 128
        // registers, 70 KB) fits beside it on the SM when both run at once (stage pipeline)
        if (threads == 128) {
            switch (split) {
            case 1: return partition_mac_multi_kernel<T, 4, 1, 128>;
            case 2: return partition_mac_multi_kernel<T, 4, 2, 128>;
            default: return partition_mac_multi_kernel<T, 4, 4, 128>;
            }
        }
        switch (split) {
        case 1: return partition_mac_multi_kernel<T, 4, 1>;
        case 2: return partition_mac_multi_kernel<T, 4, 2>;
        default: return partition_mac_multi_kernel<T, 4, 4>;
        }
    }
    switch (split) {
    case 1: return partition_mac_multi_kernel<float, 4, 1>;
    case 2: return partition_mac_multi_kernel<float, 4, 2>;
    case 4: return partition_mac_multi_kernel<float, 4, 4>;
    case 8: return partition_mac_multi_kernel<float, 4, 8>;
    default: return partition_mac_multi_kernel<float, 4, 16>;
    }
}
inline int mac_quad_split(int realsize, int split) { return realsize == 8 && split > 4 ? 4 : split; }
template <class T> inline mac_kernel_t mac_pair_kernel_for_split(int split)
{
    constexpr int U = sizeof(T) == 8 ? 2 : 4;
    switch (split) {
    case 1: return partition_mac_pair_kernel<T, 1, U>;
    case 2: return partition_mac_pair_kernel<T, 2, U>;
    case 4: return partition_mac_pair_kernel<T, 4, U>;
    case 8: return partition_mac_pair_kernel<T, 8, 2>;
    default: return partition_mac_pair_kernel<T, 16, 1>;
    }
}
template <class T> inline mac_kernel_t mac_kernel_for_split(int split)
{
    switch (split) {
    case 1: return partition_mac_kernel<T, 1, 4>;
    case 2: return partition_mac_kernel<T, 2, 4>;
    case 4: return partition_mac_kernel<T, 4, 4>;
    case 8: return partition_mac_kernel<T, 8, 4>;
    case 16: return partition_mac_kernel<T, 16, 2>;
    default: return partition_mac_kernel<T, 32, 1>;
    }
}
#endif

} // namespace bfir
