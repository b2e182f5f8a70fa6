/*
 * bfir_b200.h -- C ABI of libbfir_b200.so, the B200 (sm_100a) drop-in for the BruteFIR convolver path
 * of vsu/foo-dsp-bfir. Plain pointers and sizes only; no CUDA, C++ or torch types cross this boundary.
 *
 * Two nested surfaces are exported, mirroring the reference (citations relative to the reference tree):
 *
 *   ENGINE    bfir_*        replaces class `brutefir`            brutefir/brutefir.hpp:15-52
 *             (what foo_dsp_bfir/foo_dsp_bfir.cpp:279-332 constructs and calls once per block)
 *   CONVOLVER bfir_conv_*   replaces class `fftw_convolver`      brutefir/fftw_convolver.hpp:28-166
 *             (per-function entry points on opaque cbufs; here the cbufs are DEVICE buffers obtained
 *              from bfir_conv_alloc, because the reference's callers own, memset and memcpy them)
 *
 * Every function returns BFIR_OK (0) or a negative code; nothing throws across the boundary (the
 * reference's bare `throw;` paths, brutefir/fftw_convolver.cpp:67,73,92, become BFIR_ERR_INVALID).
 * There is NO CPU fallback: without a CUDA device every create call fails with BFIR_ERR_CUDA.
 *
 * Threading: like the reference, one thread per handle; different handles may be used from different
 * threads concurrently (each owns a CUDA stream, there is no process-global mutable state).
 */
#ifndef BFIR_B200_H
#define BFIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define BFIR_OK 0
#define BFIR_ERR_NONFINITE (-1)   /* brutefir::run returned -1 (NaN/Inf in the system), brutefir.cpp:316-321 */
#define BFIR_ERR_COEFF (-2)       /* set_coeff returned -2 (NaN/Inf among coefficients), brutefir.cpp:219-224 */
#define BFIR_ERR_NOT_READY (-3)   /* run before set_coeff / channel without coefficients */
#define BFIR_ERR_INVALID (-4)     /* invalid argument or unsupported size */
#define BFIR_ERR_CUDA (-5)        /* CUDA runtime failure (see bfir_last_error) */

/* sample formats: brutefir/global.h:24-37 */
#define BFIR_SAMPLE_FORMAT_S8 1
#define BFIR_SAMPLE_FORMAT_S16_LE 2
#define BFIR_SAMPLE_FORMAT_S16_BE 3
#define BFIR_SAMPLE_FORMAT_S24_LE 4
#define BFIR_SAMPLE_FORMAT_S24_BE 5
#define BFIR_SAMPLE_FORMAT_S32_LE 6
#define BFIR_SAMPLE_FORMAT_S32_BE 7
#define BFIR_SAMPLE_FORMAT_FLOAT_LE 8
#define BFIR_SAMPLE_FORMAT_FLOAT_BE 9
#define BFIR_SAMPLE_FORMAT_FLOAT64_LE 10
#define BFIR_SAMPLE_FORMAT_FLOAT64_BE 11

/* mix modes: brutefir/fftw_convolver.hpp:14-16 (INPUT_ADD is unimplemented in the reference too) */
#define BFIR_MIXMODE_INPUT 1
#define BFIR_MIXMODE_INPUT_ADD 2
#define BFIR_MIXMODE_OUTPUT 3

/* struct bfoverflow_t, brutefir/global.h:96-102 */
typedef struct bfir_overflow_t {
    unsigned int n_overflows;
    int32_t intlargest;
    double largest;
    double max;
} bfir_overflow_t;

typedef struct bfir_engine bfir_engine;
typedef struct bfir_conv bfir_conv;
typedef struct bfir_eq bfir_eq;

/* ------------------------------------------------------------------------------------------------
 * ENGINE -- class brutefir
 * ---------------------------------------------------------------------------------------------- */

/* Extended construction parameters. The first eight fields are the reference constructor's
 * (brutefir/brutefir.hpp:18-25); the rest have no reference counterpart and default to
 * {1, -1, 0, 0, 0, 0, 0} through bfir_create. */
typedef struct bfir_config_t {
    int filter_length;   /* block length L, power of two, 16..32768 in both precisions (transforms of 32..65536 points);
                            the largest sizes run as transforms split over two CTAs (L 32768 float, 16384 double) or
                            over a cluster of four (L 32768 double) */
    int filter_blocks;   /* partitions P >= 1 */
    int realsize;        /* 4 or 8 */
    int channels;        /* channels per stream; no BF_MAXCHANNELS limit (global.h:21) */
    int in_format;       /* BFIR_SAMPLE_FORMAT_* */
    int out_format;
    int sampling_rate;
    int apply_dither;    /* integer output formats only (fftw_convolver.cpp:421,444) */
    int n_streams;       /* independent streams batched in one engine; buffers are [stream][frame][channel] */
    int device;          /* CUDA device ordinal, -1 = current device */
    int part_begin;      /* partition shard [part_begin, part_begin+part_count) convolved by this engine; */
    int part_count;      /* 0 = all. Sharded engines produce PARTIAL spectra, see bfir_run_partial_device */
    int n_groups;        /* channel-group pipelining: whole streams are dealt to this many CUDA streams so that
                            H2D / kernels / D2H of different groups overlap; 0 = choose, 1 = off, max 8 */
    int xbar_inputs;     /* crossbar (BASELINE configs[4]): > 0 makes `channels` the number of FILTERS per stream; */
    int xbar_outputs;    /* the raw input has xbar_inputs channels, the raw output xbar_outputs channels, joined
                            to the filters by the gain matrices of bfir_set_crossbar. 0 / 0 = the reference's
                            diagonal routing (input n -> filter n -> output n, brutefir.cpp:213-216) */
} bfir_config_t;

/* brutefir::brutefir (brutefir.cpp:21-44). On failure *out is NULL. */
int bfir_create(bfir_engine **out, int filter_length, int filter_blocks, int realsize, int channels,
                int in_format, int out_format, int sampling_rate, int apply_dither);
int bfir_create_ex(bfir_engine **out, const bfir_config_t *cfg);
/* brutefir::~brutefir (brutefir.cpp:47-62) */
void bfir_destroy(bfir_engine *e);
/* brutefir::is_initialized (brutefir.cpp:68-72): 1 once set_coeff succeeded */
int bfir_is_initialized(const bfir_engine *e);

/* brutefir::set_coeff(void **coeffs, n_coeffs, length, coeff_blocks, scale) (brutefir.cpp:180-228):
 * `coeffs[n]` = `length` HOST samples of type float (realsize 4) / double (realsize 8) for channel n;
 * n_coeffs is clamped to the channel count (all streams x channels); the arrays are copied.
 * Returns 0 or BFIR_ERR_COEFF (-2) when a scaled coefficient is NaN/Inf. */
int bfir_set_coeff(bfir_engine *e, const void *const *coeffs, int n_coeffs, int length, int coeff_blocks, double scale);

/* Coefficient-set routing: filter channel c convolves with coefficient set map[c] (0 <= map[c] < channels * n_streams,
 * n = that count; map = NULL restores the identity). The original BruteFIR's struct bfcoeff_t carries the list of
 * channels a coefficient set serves (global.h:71-78); the reference always fills it with the identity
 * (brutefir.cpp:213-216) and never reads it, so this is the general form it left out: one room-correction filter
 * shared by several channels is loaded (and kept in HBM) once. Takes effect with the next block. */
int bfir_set_coeff_map(bfir_engine *e, const int *map, int n);

/* Runtime filter swap (BASELINE configs[2]; the reference has convolver_crossfade_inplace,
 * fftw_convolver.cpp:276-321, but no caller): stages a new coefficient set of the SAME geometry; the
 * next block is computed with the old and the new set and cross-faded old->new with the linear ramp of
 * crossfade_inplace (:296-305), after which the new set is current. */
int bfir_set_coeff_crossfade(bfir_engine *e, const void *const *coeffs, int n_coeffs, int length, int coeff_blocks, double scale);
/* Coefficients already on the device (e.g. rendered by bfir_equalizer_render_device): planar, channel n at
 * d_coeffs + n * channel_stride elements (stride 0 = one filter shared by all channels). */
int bfir_set_coeff_device(bfir_engine *e, const void *d_coeffs, long long channel_stride, int n_coeffs, int length,
                          int coeff_blocks, double scale, int crossfade);

/* Crossbar gains = the `scales[]` of convolver_mixnscale with n_bufs > 1 (fftw_convolver.cpp:215-229):
 * filter input f = sum_i in_gains[f * xbar_inputs + i] * input i          (MIXMODE_INPUT, n_bufs = inputs)
 * output o       = sum_f out_gains[o * channels + f] * filter output f    (MIXMODE_OUTPUT, n_bufs = filters)
 * HOST arrays of doubles, copied. The sample-format scales are applied on top, as in brutefir::run. */
int bfir_set_crossbar(bfir_engine *e, const double *in_gains, const double *out_gains);

/* Page-locked host memory for the raw blocks handed to bfir_run / bfir_run_async (the reference's
 * _aligned_realloc of m_inbuf / m_outbuf, foo_dsp_bfir.cpp:292-294). bfir_run also accepts ordinary pageable
 * memory (the driver then stages the copies itself, more slowly); bfir_run_async needs page-locked buffers.
 * Returns NULL without a CUDA device or when the allocation fails. */
void *bfir_host_alloc(size_t bytes);
void bfir_host_free(void *p);

/* brutefir::run (brutefir.cpp:245-343): inbuf/outbuf are HOST buffers holding exactly
 * n_streams * filter_length * channels interleaved samples in in_format / out_format. Synchronous:
 * returns when outbuf is filled. Returns 0, or BFIR_ERR_NONFINITE (-1) when output sample 0 of some
 * channel is NaN/Inf; in that case, as in the reference, the block counter is not advanced. */
int bfir_run(bfir_engine *e, const void *inbuf, void *outbuf);

/* Same block step on DEVICE buffers, asynchronous on the engine's stream (no error probing until
 * bfir_sync, which returns BFIR_ERR_NONFINITE if any block since the last sync hit NaN/Inf). */
int bfir_run_device(bfir_engine *e, const void *d_inbuf, void *d_outbuf);
int bfir_sync(bfir_engine *e);

/* Throughput variant of bfir_run for offline / batch callers (no reference counterpart: brutefir::run is
 * synchronous, brutefir.cpp:245-343). Queues H2D -> block step -> D2H of one block on the stream groups and
 * returns a ticket (>= 0) without waiting, so the copies and kernels of consecutive blocks overlap.
 * inbuf/outbuf are PINNED host buffers laid out as for bfir_run; they must stay untouched until
 * bfir_wait(ticket) (or bfir_sync) has returned. At most 16 steps are in flight: a seventeenth call first waits
 * for the oldest. bfir_wait returns when the step with that ticket and all earlier ones have delivered
 * their output; BFIR_ERR_NONFINITE if a NaN/Inf probe fired in any block since the last wait/sync (all
 * queued work is drained first; unlike bfir_run the block counter is not rolled back). Any other call
 * on the engine joins the queued steps before it does its own work. Returns an error code < 0 instead
 * of a ticket on failure. */
long long bfir_run_async(bfir_engine *e, const void *inbuf, void *outbuf);
int bfir_wait(bfir_engine *e, long long ticket);
/* bfir_run_device without the join at the end of the call: the stream groups of consecutive blocks run into
 * each other, and the engine's stream does NOT see the result until bfir_join (stream-ordered, no host wait),
 * bfir_sync or any other call on the engine. A group reads only its own streams' slice of d_inbuf / writes
 * its slice of d_outbuf, in call order; the caller must not rewrite a buffer that a queued block still uses. */
int bfir_run_device_pipelined(bfir_engine *e, const void *d_inbuf, void *d_outbuf);
int bfir_join(bfir_engine *e);
/* Two consecutive blocks per call, for callers that have the next block at hand (offline rendering, the
 * pipelined paths above): both forward transforms, ONE partition-sum launch in which every coefficient spectrum
 * is read once for both blocks and every delay-line spectrum serves block t at partition i and block t+1 at
 * partition i+1 (per channel (2P + split + 2) N realsize bytes for two blocks instead of 2 (2P + 1) N realsize),
 * both inverse transforms. Same results as two bfir_run_device / bfir_run_async calls up to the summation order
 * of the partition sum. While that is not possible (fewer than filter_blocks blocks since the last reset, a
 * partition shard, a pending filter swap) the call runs the two blocks one by one.
 * bfir_run_device_pair, `pipelined`:
 *   BFIR_PAIR_JOINED (0)    like bfir_run_device: stream-ordered on the engine's stream, joined at the end of the call.
 *   BFIR_PAIR_PIPELINED (1) like bfir_run_device_pipelined: no join between calls; with one stream group everything
 *                           stays stream-ordered on the engine's stream (inputs may be produced on that stream).
 *   BFIR_PAIR_STAGED (2)    the engine's STAGE PIPELINE (one stream group; otherwise as 1): the forward
 *                           transforms of pair k+1 and the inverse transforms of pair k-1 run on two side streams beside
 *                           the partition sum of pair k. CONTRACT: the side streams are NOT ordered after later work on
 *                           the engine's stream, so d_in0 / d_in1 must be COMPLETE (their producers finished, e.g. a
 *                           stream/event synchronisation by the caller) when the call is made, and d_out0 / d_out1 are
 *                           visible to the engine's stream only after bfir_join, to the host after bfir_sync.
 * bfir_run_async_pair: pinned host buffers, returns the ticket of the SECOND block (waiting on it covers both). */
#define BFIR_PAIR_JOINED 0
#define BFIR_PAIR_PIPELINED 1
#define BFIR_PAIR_STAGED 2
int bfir_run_device_pair(bfir_engine *e, const void *d_in0, const void *d_in1, void *d_out0, void *d_out1, int pipelined);
long long bfir_run_async_pair(bfir_engine *e, const void *in0, const void *in1, void *out0, void *out1);
/* Four consecutive blocks with one partition-sum launch: a window of four delay-line spectra slides over the
 * partitions, so per channel (2P + 3 split + 4) N realsize bytes move for four blocks instead of 4 (2P + 1) N realsize
 * (split = partition slices per CTA, 1 for large batches). Both precisions (double: 214 registers, one CTA per SM).
 * Device buffers. bfir_run_device_quad is joined like bfir_run_device; bfir_run_device_quad_staged runs the four blocks
 * through the stage pipeline under the contract of BFIR_PAIR_STAGED (inputs complete at call time, outputs visible
 * after bfir_join / bfir_sync). On a partition shard, with a pending filter swap and while the delay line is still
 * filling the call runs two pairs (or four single blocks); a crossbar (bfir_set_crossbar) is part of the stages. */
int bfir_run_device_quad(bfir_engine *e, const void *const d_in[4], void *const d_out[4]);
int bfir_run_device_quad_staged(bfir_engine *e, const void *const d_in[4], void *const d_out[4]);
/* EIGHT consecutive blocks with one partition-sum launch (both precisions): per channel (2P + 15) N realsize bytes for
 * eight blocks -- 9.9 spectra per block at P = 32 against 17.75 with four blocks per launch and 65 with one. Always
 * through the stage pipeline; `staged` = 0 joins at the end of the call (like bfir_run_device), `staged` != 0 leaves the
 * pipeline open under the contract of BFIR_PAIR_STAGED. One stream group, steady state and at least ~75 000 threads of
 * work; otherwise the call runs two four-block calls. The delay line keeps filter_blocks + 15 slots for it. */
int bfir_run_device_oct(bfir_engine *e, const void *const d_in[8], void *const d_out[8], int staged);
/* Four consecutive blocks of PINNED host buffers through the stage pipeline (one stream group; otherwise, and outside
 * the steady state, two bfir_run_async_pair calls): the input copies of block b, its forward transform, the four-block
 * partition sum, the inverse transforms and the output copies run on seven streams chained by events over a ring of 12
 * staging slots (consecutive blocks alternate between two copy streams each way, so that a copy which still waits for
 * its slot never holds up the next one), and the copies of neighbouring calls overlap the kernels of this one. Returns
 * the ticket of the FOURTH block (bfir_wait on it covers all four). Keep four calls in flight to hold the link's rate
 * (16 tickets may be outstanding before a call waits for the oldest). */
long long bfir_run_async_quad(bfir_engine *e, const void *const in[4], void *const out[4]);

/* brutefir::reset (brutefir.cpp:347-367): zeroes counters and overflow statistics, NOT the buffers */
int bfir_reset(bfir_engine *e);
/* struct bfoverflow_t overflow[channel] (brutefir.hpp:126) */
int bfir_get_overflow(bfir_engine *e, int channel, bfir_overflow_t *out);
/* brutefir::check_overflows (brutefir.cpp:371-388): prints "peak: ch/count/dB" through the print
 * callback when the statistics changed since the last call; returns 1 if something was printed */
int bfir_check_overflows(bfir_engine *e);
/* dither_state_t.randtab_ptr of a channel (global.h:63-69), for parity checks of the table walk */
int bfir_get_dither_ptr(bfir_engine *e, int channel, int *out);
/* unsigned int blockcounter (brutefir.hpp:106) */
int bfir_get_blockcounter(bfir_engine *e, unsigned int *out);

/* Partition-sharded operation (one engine per GPU, no reference counterpart): run_partial_device
 * executes input FFT + this shard's partition sum and leaves the PARTIAL accumulated spectra
 * (n_channels * 2L reals, ORD layout) in the buffer returned by bfir_acc_device_ptr; after the caller
 * has summed the shards' buffers (NCCL reduce / all-reduce), run_finish_device does the output stage. */
int bfir_run_partial_device(bfir_engine *e, const void *d_inbuf);
int bfir_run_finish_device(bfir_engine *e, void *d_outbuf);
/* The same with FOUR consecutive blocks per call and no collective in between (fused reduce only, i.e. after
 * bfir_peer_setup / import; steady state: at least filter_blocks blocks since the last reset, else BFIR_ERR_NOT_READY
 * and the caller keeps using the one-block pair above): bfir_run_partial_quad_device transforms the four blocks, runs
 * ONE four-block partition sum over this rank's partitions, pushes the four partial results to their owners and
 * raises this rank's arrival flag in every peer's receive buffer; bfir_run_finish_quad_device waits (on the device)
 * until every source rank's flag has arrived, then sums and emits the own channels of the four blocks. Call them
 * back to back on every rank; a rank that never arrives is reported by bfir_sync after a time-out. */
int bfir_run_partial_quad_device(bfir_engine *e, const void *const d_in[4]);
int bfir_run_finish_quad_device(bfir_engine *e, void *const d_out[4]);
/* Both halves in ONE call through the engine's stage pipeline (same preconditions): the forward transforms of call
 * k+1 run on a side stream beside the four-block partition sum of call k on the engine's stream, and the arrival wait
 * + sum + output stage of call k-1 on a third. The cross-rank protocol stays collective-free: a rank pushes call k only
 * after it has seen every peer's flag of call k-1, and raises its flag of call k only after its own output stage of
 * call k-1 has read the receive-buffer phases that call k+1 will overwrite. CONTRACT as BFIR_PAIR_STAGED: the four
 * input blocks are complete when the call is made; the compact own-channel output blocks are visible to the engine's
 * stream after bfir_join, to the host after bfir_sync. Every rank makes the same sequence of calls.
 * BFIR_SHARD_INPUTS=1 (read by bfir_peer_setup; crossbar engines): the input stage is sharded as well -- a rank
 * transforms only its own ceil(inputs / world) input channels and stores the spectra into every peer's input region
 * (a second flag array orders it), then applies the whole input crossbar. Off by default: measured slower than the
 * redundant input stage on 2 and on 8 GPUs, with one launch per block and with multi-block launches. */
int bfir_run_shard_quad_staged(bfir_engine *e, const void *const d_in[4], void *const d_out[4]);
/* The same with EIGHT blocks per call (one eight-block partition sum over the rank's partitions; shards too small for
 * that kernel run two four-block calls). */
int bfir_run_shard_oct_staged(bfir_engine *e, const void *const d_in[8], void *const d_out[8]);
void *bfir_acc_device_ptr(bfir_engine *e, size_t *bytes);

/* Fused partition-shard reduce (no reference counterpart; SURVEY 8e "fused variant"). After
 * bfir_peer_setup(rank, world) + connecting every peer's receive buffer, bfir_run_partial_device stores the
 * partial spectra of each reduced channel (outputs with a crossbar, else filters) DIRECTLY into the
 * receive buffer of the rank that owns the channel (peer-mapped pointers: NVLink stores issued by the
 * producing kernel, tile by tile), and bfir_run_finish_device -- after a cross-rank barrier the caller
 * provides -- sums the `world` slots of its own channels and emits ONLY those channels, as a compact
 * interleaved block [L][own_count]. Ownership: contiguous ranges of ceil(n / world) channels.
 * Connect peers with bfir_peer_export / bfir_peer_import (CUDA IPC handle, 64 bytes, one process per GPU)
 * or bfir_peer_set_ptr (same process). One stream, no dither, world <= 8. */
int bfir_peer_setup(bfir_engine *e, int rank, int world);
int bfir_peer_export(bfir_engine *e, void *handle64);
int bfir_peer_import(bfir_engine *e, int peer_rank, const void *handle64);
int bfir_peer_set_ptr(bfir_engine *e, int peer_rank, void *d_recv);
void *bfir_peer_recv_ptr(bfir_engine *e);
int bfir_peer_own_channels(bfir_engine *e, int *first, int *count);

/* Change the number of channel groups (see bfir_config_t.n_groups); synchronises the engine. */
int bfir_set_groups(bfir_engine *e, int n_groups);
int bfir_get_groups(bfir_engine *e);
/* partition slices per CTA the engine chose for its partition-sum kernels (reporting only) */
int bfir_get_mac_split(bfir_engine *e);
int bfir_get_quad_split(bfir_engine *e);   /* the same for the four-block kernel */

/* Use an existing CUDA stream (a cudaStream_t passed as void*) instead of the engine's own. */
int bfir_set_stream(bfir_engine *e, void *cuda_stream);

/* Per-kernel timing with CUDA events on the engine's stream (measurement only): after
 * bfir_set_profiling(e, n) the next n block steps are bracketed by events; bfir_get_profile returns
 * the accumulated milliseconds of {input FFT, partition MAC, output stage} and the number of blocks,
 * as of the last bfir_sync / bfir_run, optionally clearing the sums. */
int bfir_set_profiling(bfir_engine *e, int max_blocks);
int bfir_get_profile(bfir_engine *e, double ms_out[3], unsigned long long *blocks, int reset);
/* the partition-sum part of the profile, per kernel family: summed launch time and number of profiled launches of the
 * partition sums that covered `blocks_per_launch` (1, 2, 4 or 8) blocks */
int bfir_get_mac_profile(bfir_engine *e, int blocks_per_launch, double *ms_sum, unsigned long long *launches, int reset);

/* pinfo / set_print_callback (brutefir/pinfo.h:17-18) */
void bfir_set_print_callback(void (*cb)(const char *message));
/* last error text of the calling thread */
const char *bfir_last_error(void);
/* number of kernels this library has launched in the process so far (all handles) */
unsigned long long bfir_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * CONVOLVER -- class fftw_convolver. All cbuf / raw arguments are DEVICE pointers (bfir_conv_alloc),
 * all calls are asynchronous on the convolver's stream; bfir_conv_download / bfir_conv_sync wait.
 * ---------------------------------------------------------------------------------------------- */

/* fftw_convolver::fftw_convolver(length, realsize, dither) (fftw_convolver.cpp:51-138). The dither
 * object of the reference (dither.cpp:21-110) is built inside for `n_dither_channels` channels. */
int bfir_conv_create(bfir_conv **out, int length, int realsize, int n_dither_channels, int sampling_rate);
void bfir_conv_destroy(bfir_conv *c);
/* convolver_cbufsize (fftw_convolver.cpp:469-472): 2 * length * realsize bytes */
int bfir_conv_cbufsize(const bfir_conv *c);

void *bfir_conv_alloc(bfir_conv *c, size_t bytes);   /* zero-filled device memory */
void bfir_conv_free(bfir_conv *c, void *d_ptr);
int bfir_conv_upload(bfir_conv *c, void *d_dst, const void *h_src, size_t bytes);
int bfir_conv_download(bfir_conv *c, void *h_dst, const void *d_src, size_t bytes);
int bfir_conv_sync(bfir_conv *c);

/* convolver_raw2cbuf (fftw_convolver.cpp:157-185): bf = {format, byte_offset, sample_spacing} */
int bfir_conv_raw2cbuf(bfir_conv *c, const void *d_rawbuf, void *cbuf, void *next_cbuf, int format,
                       int byte_offset, int sample_spacing);
/* convolver_time2freq (fftw_convolver.cpp:188-212): T -> HC, in == out allowed */
int bfir_conv_time2freq(bfir_conv *c, const void *input_cbuf, void *output_cbuf);
/* convolver_mixnscale (fftw_convolver.cpp:215-229): input_cbufs = HOST array of n_bufs device cbufs;
 * output must not alias an input; n_bufs <= 64 */
int bfir_conv_mixnscale(bfir_conv *c, void *const *input_cbufs, void *output_cbuf, const double *scales,
                        int n_bufs, int mixmode);
/* convolver_convolve_inplace / convolve / convolve_add (fftw_convolver.cpp:232-273) on ORD cbufs */
int bfir_conv_convolve_inplace(bfir_conv *c, void *cbuf, const void *coeffs);
int bfir_conv_convolve(bfir_conv *c, const void *input_cbuf, const void *coeffs, void *output_cbuf);
int bfir_conv_convolve_add(bfir_conv *c, const void *input_cbuf, const void *coeffs, void *output_cbuf);
/* convolver_crossfade_inplace (fftw_convolver.cpp:276-321), float-branch algorithm in both precisions.
 * input_cbuf (ORD, new filter) becomes the cross-faded spectrum; crossfade_cbuf (ORD, old filter) and
 * buffer_cbuf are destroyed. */
int bfir_conv_crossfade_inplace(bfir_conv *c, void *input_cbuf, void *crossfade_cbuf, void *buffer_cbuf);
/* convolver_dirac_convolve[_inplace] (fftw_convolver.cpp:324-348) on HC cbufs */
int bfir_conv_dirac_convolve(bfir_conv *c, const void *input_cbuf, void *output_cbuf);
int bfir_conv_dirac_convolve_inplace(bfir_conv *c, void *cbuf);
/* convolver_freq2time (fftw_convolver.cpp:351-375): HC -> time, in == out allowed */
int bfir_conv_freq2time(bfir_conv *c, const void *input_cbuf, void *output_cbuf);
/* convolver_convolve_eval (fftw_convolver.cpp:378-403): buffer_cbuf is 1.5 cbufs, zeroed before first use */
int bfir_conv_convolve_eval(bfir_conv *c, const void *input_cbuf, void *buffer_cbuf, void *output_cbuf);
/* Small one-shot convolver, td_conv_t (fftw_convolver.hpp:18-26): convolver_td_block_length / _td_new /
 * _td_convolve (fftw_convolver.cpp:698-777) with convolve_inplace_ordered (:820-856). `h_coeffs` are
 * n_coeffs HOST samples of the convolver's precision; the overlap block is a DEVICE buffer of
 * 2 * blocklen reals, transformed in place: R2HC, complex product with the spectrum of
 * [0_blocklen | coeffs | 0] / (2 blocklen) on the plain half-complex layout, HC2R. n_coeffs == 1, where
 * the reference shifts by -1 (log2.h:38-43), is refused; blocklen must be a supported transform size
 * (>= 16). The reference never frees a td_conv_t; bfir_conv_td_free does. */
typedef struct bfir_td_conv bfir_td_conv;
int bfir_conv_td_block_length(int n_coeffs);
int bfir_conv_td_new(bfir_conv *c, bfir_td_conv **out, const void *h_coeffs, int n_coeffs);
int bfir_conv_td_blocklen(const bfir_td_conv *tdc);
const void *bfir_conv_td_coeffs(const bfir_td_conv *tdc); /* device pointer, 2 * blocklen reals (HC) */
int bfir_conv_td_convolve(bfir_conv *c, bfir_td_conv *tdc, void *d_overlap_block);
void bfir_conv_td_free(bfir_td_conv *tdc);
/* convolver_cbuf2raw (fftw_convolver.cpp:406-466): `overflow` is a HOST struct, read and updated;
 * synchronous. dither_channel selects the dither_state_t. */
int bfir_conv_cbuf2raw(bfir_conv *c, const void *cbuf, void *d_outbuf, int format, int byte_offset,
                       int sample_spacing, int apply_dither, int dither_channel, bfir_overflow_t *overflow);
/* convolver_coeffs2cbuf (fftw_convolver.cpp:475-537): `coeffs` are n_coeffs HOST samples; d_dest is a
 * device cbuf (the reference's optional_dest). Synchronous; BFIR_ERR_COEFF on NaN/Inf. */
int bfir_conv_coeffs2cbuf(bfir_conv *c, const void *coeffs, int n_coeffs, double scale, void *d_dest);
/* convolver_runtime_coeffs2cbuf (fftw_convolver.cpp:540-567): d_src = length device samples */
int bfir_conv_runtime_coeffs2cbuf(bfir_conv *c, const void *d_src, void *d_dest);
/* dither object inspection for parity: table bytes, map [-256..255] (512 entries of realsize), pointers */
int bfir_conv_dither_table_size(bfir_conv *c);
int bfir_conv_dither_table(bfir_conv *c, int8_t *h_out, int n);
int bfir_conv_dither_map(bfir_conv *c, void *h_out);
int bfir_conv_dither_ptr(bfir_conv *c, int channel);

/* ------------------------------------------------------------------------------------------------
 * EQUALIZER -- class equalizer (brutefir/equalizer.hpp:66-115): 31-band ISO 1/3-octave magnitude/phase
 * -> linear-phase FIR of taps/2 samples, taps = block_length * n_blocks (a power of two, 32 .. 2^28),
 * identical for every channel. generate() (equalizer.cpp:87-140) + render_f/d (:212-394) run on the
 * device; the reference's WAV cache is out of scope.
 * ---------------------------------------------------------------------------------------------- */
int bfir_eq_create(bfir_eq **out, int block_length, int n_blocks, int realsize, int sampling_rate);
void bfir_eq_destroy(bfir_eq *q);
int bfir_eq_taps(const bfir_eq *q);
/* n_bands <= 31 (freq ascending, mag in dB, phase as the reference takes it); h_out = taps/2 samples of realsize */
int bfir_eq_render(bfir_eq *q, int n_bands, const double *freq, const double *mag, const double *phase, void *h_out);
/* same, result left on the device: returns a device pointer to taps/2 samples (valid until the next render /
 * destroy) for bfir_set_coeff_device(..., channel_stride 0, ...), or NULL on error */
const void *bfir_eq_render_device(bfir_eq *q, int n_bands, const double *freq, const double *mag, const double *phase);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BFIR_B200_H */
