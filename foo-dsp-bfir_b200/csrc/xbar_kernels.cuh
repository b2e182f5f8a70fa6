// Crossbar mixing in the frequency domain: convolver_mixnscale with n_bufs > 1
// (reference brutefir/fftw_convolver.cpp:215-229, 908-1158, 1187-1421), i.e. a small dense real gain
// matrix applied to whole spectra:   out[o][j] = sum_i gain[o][i] * in[i][j],  j = 0..N-1.
//
// The engine uses it twice per block when a crossbar is configured (BASELINE configs[4], "32x32
// mixnscale crossbar"): inputs -> filter inputs (into the delay-line slot) and filter outputs ->
// outputs. Both sides are in ORD layout, so unlike the stand-alone entry point no reordering is
// involved. The gains are real scalars, so this is n_out*n_in FFMAs per real -- a bandwidth-trivial
// pass (reads n_in*N, writes n_out*N); tensor cores would need the gains split into several TF32
// terms to stay inside the 1e-5 budget and buy nothing at these sizes.
#pragma once
#include "rfft_kernels.cuh"

namespace bfir {

#define BFIR_MAX_XBAR 64

struct XbarArgs {
    const void *in;            // [streams * n_in][in_stride] reals
    void *out;                 // [streams * n_out][out_stride] reals (+ slot * slot_stride when state != NULL)
    long long in_stride, out_stride, slot_stride;
    const void *gains;         // device, [n_out][n_in] of T, row-major
    int n_in, n_out, N, n_streams;
    const EngineState *state;  // slot = (blockcounter + slot_offset) % n_slots
    int n_slots;
    int slot_offset;           // 1: the second block of a pair
    int use_abs_block;         // 1: the block index is abs_block, given by the host (stage pipeline: the device counter lags behind)
    unsigned int abs_block;
    int n_parts;               // filter partitions (procblocks cap)
    int *procblocks;           // [streams * n_out]: incremented here for the filter channels (brutefir.cpp:265-268)
    unsigned char *pb_inc;
    int stream_base;           // first stream of this launch (channel groups)
    PeerPush push;             // enabled: output rows go to the owner rank's receive buffer (fused reduce)
    const EngineState *push_state; // block parity of the receive buffer (one-block calls)
    int push_phase;            // >= 0: explicit receive-buffer phase (four-block calls), else the parity of push_state
    // several consecutive blocks in ONE launch (grid.z = n_multi): block z reads in_multi[z], writes out_multi[z] (NULL: `out`,
    // e.g. the delay line, where the slot follows from abs_block + z), pushes into phase push_phase + z
    int n_multi;
    const void *in_multi[8];
    void *out_multi[8];
};

#ifdef __CUDACC__
// grid: (ceil(N / 256), streams), 256 threads; one thread = one real index j of one stream
template <class T, int MAXI>
__global__ void __launch_bounds__(256) xbar_mix_kernel(const XbarArgs a)
{
    extern __shared__ __align__(16) unsigned char xbar_smem[];
    T *g = reinterpret_cast<T *>(xbar_smem);
    for (int k = threadIdx.x; k < a.n_in * a.n_out; k += blockDim.x) g[k] = ((const T *)a.gains)[k];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y + a.stream_base;
    if (a.procblocks != NULL && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x < a.n_out) {
        const int ch = s * a.n_out + threadIdx.x;
        const int pb = a.procblocks[ch];
        const int want = a.n_multi > 0 ? a.n_multi : 1;            // (multi-block launches run in the steady state: no increment left)
        const int add = min(want, max(a.n_parts - pb, 0));
        if (add > 0) a.procblocks[ch] = pb + add;
        a.pb_inc[ch] = add > 0 ? 1 : 0;
    }
    if (j >= a.N) return;
    const int z = a.n_multi > 0 ? (int)blockIdx.z : 0;
    const T *in = (const T *)(a.n_multi > 0 ? a.in_multi[z] : a.in) + (long long)s * a.n_in * a.in_stride + j;
    T x[MAXI];
#pragma unroll
    for (int i = 0; i < MAXI; i++) x[i] = i < a.n_in ? in[(long long)i * a.in_stride] : (T)0;
    long long off = (long long)s * a.n_out * a.out_stride + j;
    if (a.state != NULL) {
        const unsigned int blk = (a.use_abs_block ? a.abs_block : a.state->blockcounter + (unsigned int)a.slot_offset) + (unsigned int)z;
        off += (long long)(blk % (unsigned int)a.n_slots) * a.slot_stride;
    }
    T *out = (T *)((a.n_multi > 0 && a.out_multi[z] != NULL) ? a.out_multi[z] : a.out) + off;
    // the gains are read from shared memory by every thread of the warp at once (broadcast): one wavefront per load, so the
    // loop is bound by the number of loads -- 16-byte loads (4 float / 2 double gains) when the rows allow it
    const bool vec = (a.n_in & 3) == 0;
    for (int o = 0; o < a.n_out; o++) {
        const T *row = g + o * a.n_in;
        T acc = (T)0;
        if (vec) {
            if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int i = 0; i < MAXI; i += 4) if (i < a.n_in) {
                    const float4 w = *reinterpret_cast<const float4 *>(row + i);
                    acc = fma(w.x, x[i], acc); acc = fma(w.y, x[i + 1], acc); acc = fma(w.z, x[i + 2], acc); acc = fma(w.w, x[i + 3], acc);
                }
            } else {
#pragma unroll
                for (int i = 0; i < MAXI; i += 2) if (i < a.n_in) {
                    const double2 w = *reinterpret_cast<const double2 *>(row + i);
                    acc = fma(w.x, x[i], acc); acc = fma(w.y, x[i + 1], acc);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < MAXI; i++) if (i < a.n_in) acc = fma(row[i], x[i], acc);
        }
        if (a.push.enabled) peer_dst<T>(a.push, s * a.n_out + o, a.N, a.push_phase >= 0 ? (unsigned int)(a.push_phase + z) : (a.push_state->blockcounter & 1u))[j] = acc;
        else out[(long long)o * a.out_stride] = acc;
    }
}

// owner side of the fused reduce: sum the `world` source slots of each owned channel in rank order
// (deterministic) into the local spectrum buffer at the channel's absolute position
// grid.z > 1: consecutive blocks of a multi-block call, block z from phase + z into dst + z * dst_block_stride
template <class T>
__global__ void __launch_bounds__(256) peer_sum_kernel(const T *recv, T *dst, const EngineState *state, int world, int cpr, int N, int ch_first, int n_own, int phase,
                                                       long long dst_block_stride = 0)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int local = blockIdx.y;
    if (j >= N || local >= n_own) return;
    dst += (long long)blockIdx.z * dst_block_stride;
    const unsigned int parity = phase >= 0 ? (unsigned int)phase + blockIdx.z : (state->blockcounter & 1u);
    T acc = (T)0;
    for (int s = 0; s < world; s++) acc += recv[(((long long)parity * world + s) * cpr + local) * N + j];
    dst[(long long)(ch_first + local) * N + j] = acc;
}

// Cross-rank ordering of the four-block calls without a collective: after its pushes (earlier kernels on the same
// stream) a rank raises its arrival flag in every peer's receive buffer -- fence, then a system-scope store over
// NVLink -- and the owner's output stage starts with a kernel that waits until every source rank's flag has reached
// the call's epoch. Replaces the one-element NCCL all-reduce of the one-block calls (tens of microseconds at 8 ranks).
// A rank that never arrives trips the time-out (about two seconds) and is reported by bfir_sync, not by a hang.
// flag_off: p.flag_offset (partial output spectra) or p.flag_in_offset (input spectra of the sharded input stage)
static __global__ void peer_signal_kernel(const PeerPush p, unsigned int epoch, long long flag_off)
{
    const int q = threadIdx.x;
    if (q >= p.world) return;
    __threadfence_system();
    volatile unsigned int *f = (volatile unsigned int *)((char *)p.recv[q] + flag_off) + p.self;
    *f = epoch;
    __threadfence_system();
}
static __global__ void peer_wait_kernel(const PeerPush p, unsigned int epoch, int *timed_out, long long flag_off)
{
    const int s = threadIdx.x;
    if (s >= p.world) return;
    volatile unsigned int *f = (volatile unsigned int *)((char *)p.recv[p.self] + flag_off) + s;
    const long long t0 = clock64();
    while ((int)(*f - epoch) < 0) {
        if (clock64() - t0 > 4000000000LL) { *timed_out = 1; break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

// Sharded input stage: this rank's input spectra (rows [row0, row0 + nrows) of phase `phase`, N reals each, already in
// the local input region) -> the same rows of every peer's input region, 16-byte stores over NVLink.
// grid (ceil(row_bytes / 16 / 256), nrows * blocks, world - 1): blockIdx.y = block * nrows + row, block b in phase phase + b
static __global__ void __launch_bounds__(256) peer_bcast_kernel(const PeerPush p, int phase, int row0, long long row_bytes, int nrows)
{
    int q = blockIdx.z;
    if (q >= p.self) q++;                               // the z-th OTHER rank
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 16 >= row_bytes) return;
    const int b = blockIdx.y / nrows, k = blockIdx.y - b * nrows;
    const long long off = p.xin_offset + ((long long)(phase + b) * p.n_inputs + row0 + k) * row_bytes;
    const uint4 v = ((const uint4 *)((const char *)p.recv[p.self] + off))[i];
    ((uint4 *)((char *)p.recv[q] + off))[i] = v;
}

// nb raw interleaved blocks -> planar rows plan[z][channel][L] (the several-blocks-per-launch forward transforms read
// those instead of the interleaved block, which makes the blocks of a call independent of each other).
// grid (ceil(L/256), channels, nb)
struct PlanarArgs {
    const void *raw[8];
    void *plan;              // [nb][n_ch][L] reals
    int L, n_ch, fmt, nb;
};
template <class T>
static __global__ void __launch_bounds__(256) raw_to_planar_kernel(const PlanarArgs a)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, z = blockIdx.z;
    if (f >= a.L) return;
    const int bytes = fmt_bytes(a.fmt);
    const T x = load_raw<T>((const uint8_t *)a.raw[z] + ((long long)f * a.n_ch + c) * bytes, a.fmt);
    ((T *)a.plan)[((long long)z * a.n_ch + c) * a.L + f] = x;
}

typedef void (*xbar_kernel_t)(const XbarArgs);
template <class T> inline xbar_kernel_t xbar_kernel_for(int n_in)
{
    if (n_in <= 4) return xbar_mix_kernel<T, 4>;
    if (n_in <= 8) return xbar_mix_kernel<T, 8>;
    if (n_in <= 16) return xbar_mix_kernel<T, 16>;
    if (n_in <= 32) return xbar_mix_kernel<T, 32>;
    return xbar_mix_kernel<T, 64>;
}
#endif

} // namespace bfir
