// The engine: device-resident counterpart of the reference's `brutefir` class
// (brutefir/brutefir.hpp:15-130, brutefir/brutefir.cpp). One block step = three kernels on one stream:
//
//   rfft_forward_kernel   raw2cbuf + time2freq + mixnscale(INPUT)  -> FDL slot (blockcounter % P)
//   partition_mac_kernel  convolve + (P-1) x convolve_add          -> accumulated spectrum
//   rfft_inverse_kernel   mixnscale(OUTPUT) + freq2time + probe + cbuf2raw -> interleaved raw output
//   (+ dither_kernel when integer output is dithered)
//
// All state the reference keeps in host members (blockcounter, procblocks, dither_state, overflow;
// brutefir.hpp:104-127) lives in device memory, so a block step needs no host-side bookkeeping.
#include "common.hpp"
#include "xbar_kernels.cuh"

namespace bfir {

struct Engine {
    bfir_config_t cfg;
    int L = 0, N = 0, P = 0, C = 0, S = 0, Ct = 0, rs = 0, log2m = 0;
    int Pslots = 0;                 // delay-line slots per channel, P + 15 (see init)
    // crossbar: Ci inputs and Co outputs per stream around the C filters (== C without a crossbar)
    int Ci = 0, Co = 0, Cit = 0, Cot = 0;
    bool xbar = false, xbar_set = false;
    void *xin = nullptr, *yacc = nullptr, *gains_in = nullptr, *gains_out = nullptr;
    // filter swap with crossfade (configs[2]): staged coefficient set, second accumulator, two time buffers
    void *coeffs_next = nullptr, *acc2 = nullptr, *tbuf = nullptr;
    bool xfade_pending = false;
    // fused partition-shard reduce: peer-mapped receive buffers (PeerPush), this rank's own channel range
    PeerPush peer = {};
    void *recv = nullptr;
    void *peer_opened[BFIR_MAX_PEERS] = {};
    int own_first = 0, own_count = 0;
    int peer_phase = -1;            // receive-buffer phase back_group sums (-1: block parity, one-block calls)
    unsigned int peer_epoch = 0;    // four-block calls made (arrival-flag value)
    unsigned int peer_in_epoch = 0; // staged four-block calls with the sharded input stage (input-flag value)
    int own_in_first = 0, own_in_count = 0;   // inputs this rank transforms in the sharded input stage
    bool shard_inputs = false;      // BFIR_SHARD_INPUTS=1: sharded input stage (measured slower: 0.129 against 0.112 ms per block on 8 GPUs)
    // BFIR_SHARD_TIMING=1 (diagnosis): CUDA-event marks at the stage boundaries of every staged shard call, summarised on stderr by destroy()
    bool shard_timing = false;
    std::vector<cudaEvent_t> st_marks;   // 8 per call: fwd begin, fwd end, sum begin, sum end, pushes end, flag out, flags in, output stage end
    void st_mark(cudaStream_t st) { if (!shard_timing || st_marks.size() >= 8 * 512) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); st_marks.push_back(e); }
    void shard_report();
    bool peer_quad_pending = false; // run_partial_quad has been queued, run_finish_quad has not
    int *d_peer_timeout = nullptr;  // set by peer_wait_kernel when a source rank never arrived
    int run_partial_quad(const void *const d_in[4]);
    int run_finish_quad(void *const d_out[4]);
    int shard_blocks_staged(int nb, const void *const *d_in, void *const *d_out);
    // multi-block launches of the stage pipeline's output stage (BFIR_BATCH_STAGES=0: one launch per block as before)
    int mac_persist_sms = 0;        // BFIR_MAC_SMS: > 0 = the eight-block partition sum as a persistent grid alone on that many SMs
    bool batch_stages = true;
    void *out8 = nullptr;           // [8][max(Cot, Ct)][N]: per-block spectra between the output crossbar / slot sum and the inverse transforms
    size_t out8_stride = 0;         // bytes per block
    int back_blocks(int nb, void *const *acc_in, void *const *d_out, cudaStream_t st);
    void *plan8 = nullptr, *xin8 = nullptr;   // [8][Cit][L] planar raw blocks, [8][Cit][N] input spectra (crossbar engines, front_blocks)
    int front_blocks(int nb, const void *const *d_in, cudaStream_t st);
    void *acc_oct[4] = {};          // accumulated spectra of blocks 5 .. 8 of an eight-block shard call
    cudaEvent_t sp_arrived[2] = {};  // staged shard calls: every source rank's flag of call k has been seen (inverse stream)
    int peer_setup(int rank, int world);
    int peer_ready() const { if (!peer.enabled) return 1; for (int q = 0; q < peer.world; q++) if (!peer.recv[q]) return 0; return 1; }
    int part_begin = 0, part_count = 0;
    SampleFormat in_sf, out_sf;
    bool dither_on = false, initialized = false, own_stream = true;
    int device = 0;
    cudaStream_t stream = nullptr;
    void *fdl = nullptr, *coeffs = nullptr, *acc = nullptr, *prev = nullptr, *ybuf = nullptr, *tw = nullptr;
    void *d_in = nullptr, *d_out = nullptr;   // staging for the host-buffer run()
    size_t in_bytes = 0, out_bytes = 0;
    int coeff_alloc = 0;
    EngineState *state = nullptr;   // [BFIR_MAX_GROUPS], one per channel group, kept in lockstep
    // channel-group pipelining: whole streams are dealt to n_groups groups, each with its own CUDA
    // stream, so that H2D / kernels / D2H of different groups overlap (and FFT with MAC kernels)
    struct Group {
        int s0, s1, c0, c1;
        cudaStream_t stream; cudaEvent_t done;
        // bfir_run_async: the group's copies ride on their own streams so that they overlap its kernels too
        cudaStream_t h2d, d2h;
        cudaEvent_t in_ready, out_ready, copies_done;
        cudaEvent_t in_free[12], out_free[12]; // per staging slot (kStage)
    };
    Group groups[BFIR_MAX_GROUPS] = {};
    int n_groups = 1;
    cudaEvent_t fork_ev = nullptr;
    int *procblocks = nullptr, *coeff_blocks = nullptr, *nonfinite = nullptr;
    int *coeff_map = nullptr;       // [Ct] coefficient set per filter channel; nullptr = identity (bfir_set_coeff_map)
    unsigned char *pb_inc = nullptr;
    OverflowStats *stats = nullptr;
    EngineState *h_state = nullptr; // pinned, [BFIR_MAX_GROUPS]
    int *h_flag = nullptr, *d_flag = nullptr; // mapped pinned word the kernels raise on a non-finite probe
    std::vector<bfir_overflow_t> last_overflow;
    DitherTables dither;
    double ovf_max = 1.0;
    unsigned long long blocks_since_sync = 0;
    unsigned int host_blockcounter = 0;   // mirror of the device block counter (its parity selects the prev buffer)
    // the kernels of one block step, captured once per block parity and replayed (bfir_run, one group)
    cudaGraphExec_t step_graph[2] = { nullptr, nullptr };
    unsigned long long step_graph_launches[2] = { 0, 0 };
    bool graphs_enabled = true;
    int graph_flavor = 0;           // 0: forward + full partition sum + inverse; 1: forward + inverse with the head term
    // look-ahead partition sum on the synchronous bfir_run path: once a block is out, sum_{i>=1} X[t+1-i] H[i] of
    // the NEXT block is computed in the background (only old delay-line slots feed it); the next call then
    // runs forward transform -> inverse transform with the head term X[t+1] H[0] folded into its load phase
    bool lookahead_enabled = true, tail_ready = false;
    cudaEvent_t out_done = nullptr; // the output block of the current bfir_run is on the host
    // with several stream groups the look-ahead launches run one after the other on their own stream, so that
    // group 0 can start the next block while the sums of the later groups are still streaming (and the output
    // copies of the groups come out staggered instead of all at once)
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t tail_done[BFIR_MAX_GROUPS] = {};
    cudaEvent_t tail_join_ev = nullptr;
    bool tail_inflight = false;     // tail_stream holds launches the engine's stream has not been ordered after
    int join_tail();
    bool lookahead_ok() const
    {
        return lookahead_enabled && !xbar && !peer.enabled && part_begin == 0 && part_count == P && P >= 2 && !xfade_pending && pcap == 0;
    }
    int tail_group(int g, cudaStream_t st);
    void invalidate_graphs() { for (int k = 0; k < 2; k++) if (step_graph[k]) { cudaGraphExecDestroy(step_graph[k]); step_graph[k] = nullptr; } }
    int run_step_graph(bool use_tail);
    // throughput variant of run (bfir_run_async / bfir_wait): block steps are queued on the group streams
    // without joining them, so the copies and kernels of consecutive blocks overlap; a ring of events
    // (one per group) marks the completion of each queued step
    static const int kMaxInflight = 16;
    cudaEvent_t ticket_ev[kMaxInflight][BFIR_MAX_GROUPS] = {};
    long long next_ticket = 0, done_ticket = 0;    // tickets < done_ticket are known to be complete
    bool async_open = false;                        // group streams hold work the engine's stream has not joined
    bool async_copies = false;                      // ... and so do the groups' copy streams (bfir_run_async)
    int close_async();
    long long run_host_async(const void *inbuf, void *outbuf);
    int wait_ticket(long long t);
    // two blocks per call (bfir_run_device_pair / bfir_run_async_pair): both forward transforms, ONE partition-sum
    // launch that reads every coefficient spectrum once for both blocks, both inverse transforms
    void *acc_pair = nullptr;       // accumulated spectra of the second block, [Ct][N]
    // bfir_run_async: ring of staging buffers for the raw blocks (slot 0 = d_in / d_out), so that the input copies can
    // run up to kStage - 1 blocks ahead of the transforms and the output copies behind them. A slot is busy from the
    // start of its H2D to the end of its D2H (three pipeline stages), so fewer than three blocks (pairs) worth of
    // slots leaves one of the stages idle: 8 slots = 4 pairs
    static const int kStage = 12;   // three four-block calls (or six pairs) in flight
    void *stage_in[kStage] = {}, *stage_out[kStage] = {};
    unsigned long long stage_next = 0;
    int stage_count = 8;            // slots the one- and two-block host calls cycle through (BFIR_STAGE = 1 .. 12, measurement)
    int open_async_copies();
    int stage_alloc(int k);
    int fwd_block_offset = 0;       // front_group: 1 while the second block of a pair is transformed
    // stage pipeline (one group, bfir_run_device_pair(pipelined)): forward transforms of pair k+1 on their own stream under
    // the partition sum of pair k on the engine's stream, inverse transforms of pair k-1 on a third; the kernels get the
    // block index from the host because the device counter (advanced by the inverse kernels) lags behind
    cudaStream_t sp_fwd = nullptr, sp_inv = nullptr;
    cudaEvent_t sp_fwd_done[2] = {}, sp_mac_done[2] = {}, sp_inv_done[2] = {};
    void *sp_acc[2][8] = {};        // [call parity][block in call] accumulated spectra (two, four or eight blocks per call)
    unsigned long long sp_pairs = 0; // pairs queued since the pipeline was opened
    bool sp_open = false;
    bool staged_enabled = true;     // BFIR_STAGED=0 switches the stage pipeline off
    bool whole_copies = false;      // BFIR_WHOLE_COPIES=1: the pair host path moves whole blocks on one copy stream each way
    cudaStream_t stage_stream = nullptr; // front_group / back_group launch here instead of the group's stream (stage pipeline)
    bool use_abs = false;           // front_group / pair sum: pass host_blockcounter (+ offset) as the block index
    int staged_pair(const void *d_in0, const void *d_in1, void *d_out0, void *d_out1);
    int close_staged();
    const void *acc_override = nullptr; // back_group: accumulated spectra to emit instead of acc
    bool pair_ok() const
    {
        return !peer.enabled && part_begin == 0 && part_count == P && P >= 2 && !xfade_pending &&
               host_blockcounter >= (unsigned int)P;
    }
    void *acc_quad[2] = {};         // accumulated spectra of blocks 3 and 4 of a quad (single precision)
    int quad_group(int g, const void *const d_in[4], void *const d_out[4]);
    int enqueue_quad(const void *const d_in[4], void *const d_out[4], bool staged = false);
    int enqueue_oct(const void *const d_in[8], void *const d_out[8], bool staged);
    int staged_blocks(int nb, const void *const *d_in, void *const *d_out, const void *const *h_in = nullptr, void *const *h_out = nullptr);
    long long run_host_async_quad(const void *const in[4], void *const out[4]);
    cudaEvent_t q_in_ready[4] = {}, q_out_ready[4] = {};   // four-block host calls: block b has arrived / has been emitted
    int host_copy_streams = 2;      // copy streams each way of the four-block host calls (BFIR_HOST_COPY_STREAMS = 1 .. 4)
    bool host_staged_open = false;  // those streams hold copies the engine's stream has not joined
    // BFIR_COPY_TIMING=1 (diagnosis): CUDA-event brackets around every copy of the four-block host path, summarised on stderr by destroy()
    bool copy_timing = false;
    std::vector<cudaEvent_t> ct_h2d, ct_d2h;
    void copy_mark(std::vector<cudaEvent_t> &v, cudaStream_t st) { if (!copy_timing || v.size() >= 8192) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); v.push_back(e); }
    void copy_report();
    int pair_group(int g, const void *d_in0, const void *d_in1, void *d_out0, void *d_out1, cudaEvent_t *input_consumed, cudaEvent_t *output_free);
    int enqueue_pair(const void *d_in0, const void *d_in1, void *d_out0, void *d_out1, bool pipelined, bool staged = false);
    long long run_host_async_pair(const void *in0, const void *in1, void *out0, void *out1);
    int mac_split = 1;              // partition slices per CTA of the MAC kernel
    int quad_threads = 128;         // threads per CTA of the double-precision four-block kernel (BFIR_QUAD_THREADS = 128 | 256; measured 1.5 % apart)
    int quad_split = 1;             // the same for the four-block kernel (one CTA per SM in double precision: fewer slices)
    int fft_r0 = 1;                 // CTAs per transform (rfft_choose_r0)
    // optional per-kernel timing (bfir_set_profiling)
    std::vector<cudaEvent_t> pev;
    size_t pcap = 0, pidx = 0;
    double pms[3] = {0, 0, 0};
    unsigned long long pblocks = 0;
    bool prof_suppress = false;     // pair_group times the pair as a whole and silences the per-block marks inside it
    void prof(int k) { if (pidx < pcap && !prof_suppress) cudaEventRecord(pev[4 * pidx + k], stream); }
    // blocks per profiled entry (1, 2, 4, 8): lets the partition-sum time be reported per kernel family
    int prof_nb = 1;
    std::vector<int> pnb;
    double pmac_by_nb[9] = {0};
    unsigned long long pcount_by_nb[9] = {0};
    void prof_next() { if (pidx < pcap) { if (pnb.size() <= pidx) pnb.resize(pidx + 1); pnb[pidx] = prof_nb; pidx++; } }
    void prof_collect();
    void prof_free() { for (auto ev : pev) cudaEventDestroy(ev); pev.clear(); pcap = pidx = 0; }

    ~Engine() { destroy(); }
    int init(const bfir_config_t &c);
    void destroy();
    int set_coeff(const void *const *coeffs, int n_coeffs, int length, int coeff_blocks, double scale);
    int load_coeff(const void *const *h_coeffs, const void *d_src, long long d_stride, int n_coeffs, int length, int blocks, double scale, bool next);
    void finish_block();
    int set_crossbar(const double *in_gains, const double *out_gains);
    int set_groups(int n);
    cudaStream_t gstream(int g) const { return n_groups == 1 ? stream : groups[g].stream; }
    int fork();
    int join();
    int front_group(int g, const void *d_inbuf, cudaEvent_t *input_consumed = nullptr, bool skip_mac = false);
    int back_group(int g, void *d_outbuf, bool head = false);
    int enqueue_front(const void *d_inbuf);
    int enqueue_back(void *d_outbuf);
    int enqueue_block(const void *d_inbuf, void *d_outbuf);
    int enqueue_block_pipelined(const void *d_inbuf, void *d_outbuf);
    int run_host(const void *inbuf, void *outbuf);
    int sync_and_probe(bool allow_rollback = true, cudaEvent_t wait_for = nullptr);
    int reset();
    int get_overflow(int ch, bfir_overflow_t *out);
};

int Engine::init(const bfir_config_t &c)
{
    cfg = c;
    L = c.filter_length; N = 2 * L; P = c.filter_blocks; C = c.channels; S = c.n_streams > 0 ? c.n_streams : 1;
    rs = c.realsize;
    if (rs != 4 && rs != 8) { set_error("Invalid real size %d.", rs); return BFIR_ERR_INVALID; }   // fftw_convolver.cpp:64-68
    log2m = ilog2_exact(L);
    if (log2m < 0) { set_error("Invalid length %d.", L); return BFIR_ERR_INVALID; }                 // fftw_convolver.cpp:70-74
    if (!rfft_supported(rs, log2m)) { set_error("block length %d not supported for realsize %d", L, rs); return BFIR_ERR_INVALID; }
    if (P < 1 || C < 1) { set_error("No channels defined."); return BFIR_ERR_INVALID; }            // brutefir.cpp:745-749
    if ((long long)C * S > 0x7fffffffLL / 2) return BFIR_ERR_INVALID;
    Ct = C * S;
    Pslots = P + 15;
    xbar = c.xbar_inputs > 0 || c.xbar_outputs > 0;
    Ci = c.xbar_inputs > 0 ? c.xbar_inputs : C;
    Co = c.xbar_outputs > 0 ? c.xbar_outputs : C;
    if (xbar && (Ci > BFIR_MAX_XBAR || Co > BFIR_MAX_XBAR || C > BFIR_MAX_XBAR)) { set_error("crossbar larger than %d", BFIR_MAX_XBAR); return BFIR_ERR_INVALID; }
    Cit = Ci * S; Cot = Co * S;
    if (!fill_sample_format(&in_sf, c.in_format, true) || !fill_sample_format(&out_sf, c.out_format, false)) {
        set_error("invalid sample format %d / %d", c.in_format, c.out_format);
        return BFIR_ERR_INVALID;
    }
    part_begin = c.part_begin;
    part_count = c.part_count > 0 ? c.part_count : P - part_begin;
    if (part_begin < 0 || part_begin + part_count > P) { set_error("invalid partition shard"); return BFIR_ERR_INVALID; }
    dither_on = c.apply_dither && !out_sf.isfloat;                                                   // fftw_convolver.cpp:421,444
    ovf_max = out_sf.isfloat ? 1.0 : (double)(1 << ((out_sf.bytes << 3) - 1)) - 1.0;                  // brutefir.cpp:672-684

    if (c.device >= 0) BFIR_CUDA(cudaSetDevice(c.device));
    BFIR_CUDA(cudaGetDevice(&device));
    BFIR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;

    const size_t cbuf = (size_t)N * rs;
    // The reference's delay line has P slots, slot = blockcounter % P (brutefir.cpp:270,294). Here it has P + 15:
    // with P + 1 the spectrum of block t+1 can be written while block t's partition sum still needs X[t-P+1] (block
    // pairs), P + 3 covers four blocks per launch, P + 7 eight, and the stage pipeline transforms the NEXT call's (up
    // to eight) blocks while the sum of this one runs, eight more. Which slot a block lands in is not observable --
    // partition i is only read once procblocks says the slot has been written since the last reset.
    BFIR_CUDA(cudaMalloc(&fdl, cbuf * Pslots * Ct));
    BFIR_CUDA(cudaMemsetAsync(fdl, 0, cbuf * Pslots * Ct, stream));                                      // brutefir.cpp:768-769
    BFIR_CUDA(cudaMalloc(&acc, cbuf * Ct));
    BFIR_CUDA(cudaMemsetAsync(acc, 0, cbuf * Ct, stream));
    BFIR_CUDA(cudaMalloc(&prev, (size_t)2 * L * rs * Cit));                                         // input_timecbuf[n][2]
    BFIR_CUDA(cudaMemsetAsync(prev, 0, (size_t)2 * L * rs * Cit, stream));
    if (xbar) {
        BFIR_CUDA(cudaMalloc(&xin, cbuf * Cit));
        BFIR_CUDA(cudaMalloc(&yacc, cbuf * Cot));
        BFIR_CUDA(cudaMalloc(&gains_in, (size_t)C * Ci * rs));
        BFIR_CUDA(cudaMalloc(&gains_out, (size_t)Co * C * rs));
    }
    fft_r0 = rfft_choose_r0(rs, log2m, Ct);
    if (const char *env = getenv("BFIR_GRAPHS")) graphs_enabled = atoi(env) != 0;
    if (const char *env = getenv("BFIR_LOOKAHEAD")) lookahead_enabled = atoi(env) != 0;
    if (const char *env = getenv("BFIR_STAGED")) staged_enabled = atoi(env) != 0;
    if (const char *env = getenv("BFIR_COPY_TIMING")) copy_timing = atoi(env) != 0;
    if (const char *env = getenv("BFIR_SHARD_TIMING")) shard_timing = atoi(env) != 0;
    if (const char *env = getenv("BFIR_BATCH_STAGES")) batch_stages = atoi(env) != 0;
    if (const char *env = getenv("BFIR_HOST_COPY_STREAMS")) { const int v = atoi(env); if (v >= 1 && v <= 4) host_copy_streams = v; }
    if (const char *env = getenv("BFIR_MAC_SMS")) { const int v = atoi(env); if (v >= 1 && v <= 1024) mac_persist_sms = v; }
    if (const char *env = getenv("BFIR_WHOLE_COPIES")) whole_copies = atoi(env) != 0;
    if (const char *env = getenv("BFIR_STAGE")) { const int v = atoi(env); if (v >= 1 && v <= kStage) stage_count = v; }
    BFIR_CUDA(cudaEventCreateWithFlags(&out_done, cudaEventDisableTiming));
    {   // the stage pipeline's side streams; BFIR_STAGE_PRIO=1: at the highest priority, so that the block scheduler places
        // the (whole-SM) transform CTAs of the neighbouring calls before further CTAs of the running partition sum
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char *env = getenv("BFIR_STAGE_PRIO");
        const int prio = (env && atoi(env) != 0) ? prio_hi : 0;
        BFIR_CUDA(cudaStreamCreateWithPriority(&sp_fwd, cudaStreamNonBlocking, prio));
        BFIR_CUDA(cudaStreamCreateWithPriority(&sp_inv, cudaStreamNonBlocking, prio));
    }
    for (int k = 0; k < 4; k++) { BFIR_CUDA(cudaEventCreateWithFlags(&q_in_ready[k], cudaEventDisableTiming)); BFIR_CUDA(cudaEventCreateWithFlags(&q_out_ready[k], cudaEventDisableTiming)); }
    for (int k = 0; k < 2; k++)
        for (cudaEvent_t *ev : { &sp_fwd_done[k], &sp_mac_done[k], &sp_inv_done[k], &sp_arrived[k] }) BFIR_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    BFIR_CUDA(cudaStreamCreateWithFlags(&tail_stream, cudaStreamNonBlocking));
    for (int g = 0; g < BFIR_MAX_GROUPS; g++) BFIR_CUDA(cudaEventCreateWithFlags(&tail_done[g], cudaEventDisableTiming));
    BFIR_CUDA(cudaEventCreateWithFlags(&tail_join_ev, cudaEventDisableTiming));
    in_bytes = (size_t)S * L * Ci * in_sf.bytes;
    out_bytes = (size_t)S * L * Co * out_sf.bytes;
    BFIR_CUDA(cudaMalloc(&d_in, in_bytes));
    BFIR_CUDA(cudaMalloc(&d_out, out_bytes));
    BFIR_CUDA(cudaMalloc((void **)&state, sizeof(EngineState) * BFIR_MAX_GROUPS));
    BFIR_CUDA(cudaMalloc((void **)&procblocks, sizeof(int) * Ct));
    BFIR_CUDA(cudaMalloc((void **)&coeff_blocks, sizeof(int) * Ct));
    BFIR_CUDA(cudaMemsetAsync(coeff_blocks, 0, sizeof(int) * Ct, stream));
    BFIR_CUDA(cudaMalloc((void **)&pb_inc, (size_t)Ct));
    BFIR_CUDA(cudaMemsetAsync(pb_inc, 0, (size_t)Ct, stream));
    BFIR_CUDA(cudaMalloc((void **)&nonfinite, sizeof(int)));
    BFIR_CUDA(cudaMalloc((void **)&stats, sizeof(OverflowStats) * (Cot > Ct ? Cot : Ct)));
    BFIR_CUDA(cudaHostAlloc((void **)&h_state, sizeof(EngineState) * BFIR_MAX_GROUPS, cudaHostAllocDefault));
    BFIR_CUDA(cudaHostAlloc((void **)&h_flag, sizeof(int), cudaHostAllocMapped));
    *h_flag = 0;
    BFIR_CUDA(cudaHostGetDevicePointer((void **)&d_flag, h_flag, 0));
    for (int g = 0; g < BFIR_MAX_GROUPS; g++) {
        BFIR_CUDA(cudaStreamCreateWithFlags(&groups[g].stream, cudaStreamNonBlocking));
        BFIR_CUDA(cudaEventCreateWithFlags(&groups[g].done, cudaEventDisableTiming));
        BFIR_CUDA(cudaStreamCreateWithFlags(&groups[g].h2d, cudaStreamNonBlocking));
        BFIR_CUDA(cudaStreamCreateWithFlags(&groups[g].d2h, cudaStreamNonBlocking));
        for (cudaEvent_t *ev : { &groups[g].in_ready, &groups[g].out_ready, &groups[g].copies_done })
            BFIR_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        for (int k = 0; k < kStage; k++) {
            BFIR_CUDA(cudaEventCreateWithFlags(&groups[g].in_free[k], cudaEventDisableTiming));
            BFIR_CUDA(cudaEventCreateWithFlags(&groups[g].out_free[k], cudaEventDisableTiming));
        }
    }
    BFIR_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
    int rc = make_twiddles(rs, N, &tw);
    if (rc != BFIR_OK) return rc;
    if (dither_on) {
        BFIR_CUDA(cudaMalloc(&ybuf, (size_t)L * rs * Cot));
        // max_dither_table_size is 0 (memset bfconf, brutefir.cpp:31-32; passed at :712)
        rc = dither.init(Cot, c.sampling_rate, rs, 0, L);
        if (rc != BFIR_OK) return rc;
    }
    // MAC decomposition: enough CTAs for >= ~8 waves of 148 SMs x 3 resident CTAs, never more slices
    // than partitions; BFIR_MAC_SPLIT overrides (measurement only)
    {
        const long long group_threads = (long long)Ct * (N / 8);
        mac_split = 1;
        while (mac_split < 16 && mac_split * 2 <= P && group_threads * mac_split / 256 < 8LL * 148 * 3) mac_split *= 2;
        if (const char *env = getenv("BFIR_MAC_SPLIT")) {
            const int v = atoi(env);
            if ((v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) && v <= P) mac_split = v;
        }
        // four-block kernel: 214 registers in double precision (one CTA of 256 threads per SM), 128 in single (two);
        // slices only until there are ~4 waves of CTAs (measured, cfg1 x 16: 209 / 230 / 256 us at 1 / 2 / 4 slices)
        const long long resident = 148LL * (rs == 8 ? 1 : 2);
        quad_split = 1;
        while (quad_split < (rs == 8 ? 4 : 16) && quad_split * 2 <= P && group_threads * quad_split / 256 < 4 * resident) quad_split *= 2;
        if (const char *env = getenv("BFIR_QUAD_THREADS")) { const int v = atoi(env); if (v == 128 || v == 256) quad_threads = v; }
        if (const char *env = getenv("BFIR_QUAD_SPLIT")) {
            const int v = atoi(env);
            if ((v == 1 || v == 2 || v == 4 || (rs == 4 && (v == 8 || v == 16))) && v <= P) quad_split = v;
        }
    }
    last_overflow.assign(Cot, bfir_overflow_t{0, 0, 0.0, ovf_max});
    {
        // groups: only whole streams can be split (the raw buffers interleave the channels of a stream);
        // worth it once a group still fills the machine. BFIR_GROUPS overrides (measurement).
        int want = c.n_groups > 0 ? c.n_groups : 1;
        if (c.n_groups <= 0) {
            const double step_bytes = (double)Ct * (2.0 * P + 1) * N * rs;
            if (S >= 4 && step_bytes >= 256e6) want = 4;
            else if (S >= 2 && step_bytes >= 64e6) want = 2;
        }
        if (const char *env = getenv("BFIR_GROUPS")) { const int v = atoi(env); if (v >= 1) want = v; }
        rc = set_groups(want);
        if (rc != BFIR_OK) return rc;
    }
    rc = reset();
    if (rc != BFIR_OK) return rc;
    BFIR_CUDA(cudaStreamSynchronize(stream));
    return BFIR_OK;
}

void Engine::prof_collect()
{
    for (size_t i = 0; i < pidx; i++) {
        for (int k = 0; k < 3; k++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pev[4 * i + k], pev[4 * i + k + 1]) == cudaSuccess) {
                pms[k] += ms;
                if (k == 1 && i < pnb.size() && pnb[i] >= 1 && pnb[i] <= 8) { pmac_by_nb[pnb[i]] += ms; pcount_by_nb[pnb[i]]++; }
            }
        }
        pblocks++;
    }
    if (pidx > 0) { pcap -= pidx; pev.erase(pev.begin(), pev.begin() + 4 * pidx); pnb.erase(pnb.begin(), pnb.begin() + (pnb.size() < pidx ? pnb.size() : pidx)); }
    pidx = 0;
}

void Engine::shard_report()
{
    const size_t n = st_marks.size() / 8;
    if (n >= 6) {
        cudaEventSynchronize(st_marks.back());
        static const char *names[8] = { "forward begin", "forward end", "sum begin", "sum end", "pushes end", "flag out", "flags in", "output stage end" };
        double off[8] = { 0 }, period = 0.0;
        float ms = 0.f;
        const size_t first = n / 3;
        for (size_t k = first; k < n; k++) {
            for (int j = 0; j < 8; j++) { cudaEventElapsedTime(&ms, st_marks[8 * k + 2], st_marks[8 * k + j]); off[j] += ms; }
            if (k > first) { cudaEventElapsedTime(&ms, st_marks[8 * (k - 1) + 2], st_marks[8 * k + 2]); period += ms; }
        }
        fprintf(stderr, "[bfir shard timing] rank %d, %zu calls, period %.4f ms per call; offsets from the partition sum's start:", peer.self, n - first, period / (n - first - 1));
        for (int j = 0; j < 8; j++) fprintf(stderr, " %s %+.4f;", names[j], off[j] / (n - first));
        fprintf(stderr, "\n");
    }
    for (auto e : st_marks) cudaEventDestroy(e);
    st_marks.clear();
}

void Engine::copy_report()
{
    for (int dir = 0; dir < 2; dir++) {
        std::vector<cudaEvent_t> &v = dir == 0 ? ct_h2d : ct_d2h;
        const size_t n = v.size() / 2;
        if (n >= 8) {
            cudaEventSynchronize(v.back());
            const size_t first = n / 2;                       // second half: steady state
            double busy = 0.0; float ms = 0.f, span = 0.f;
            for (size_t k = first; k < n; k++) { cudaEventElapsedTime(&ms, v[2 * k], v[2 * k + 1]); busy += ms; }
            cudaEventElapsedTime(&span, v[2 * first], v[2 * n - 1]);
            fprintf(stderr, "[bfir copy timing] %s: %zu copies, mean copy %.4f ms, span per copy %.4f ms, busy fraction %.3f\n",
                    dir == 0 ? "H2D" : "D2H", n - first, busy / (n - first), span / (n - first), busy / span);
        }
        for (auto e : v) cudaEventDestroy(e);
        v.clear();
    }
}

void Engine::destroy()
{
    if (copy_timing) copy_report();
    if (shard_timing) shard_report();
    prof_free();
    if (stream) cudaStreamSynchronize(stream);
    invalidate_graphs();
    for (int g = 0; g < BFIR_MAX_GROUPS; g++) {
        if (groups[g].stream) { cudaStreamSynchronize(groups[g].stream); cudaStreamDestroy(groups[g].stream); groups[g].stream = nullptr; }
        if (groups[g].done) { cudaEventDestroy(groups[g].done); groups[g].done = nullptr; }
        for (cudaStream_t *st : { &groups[g].h2d, &groups[g].d2h }) if (*st) { cudaStreamSynchronize(*st); cudaStreamDestroy(*st); *st = nullptr; }
        for (cudaEvent_t *ev : { &groups[g].in_ready, &groups[g].out_ready, &groups[g].copies_done })
            if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
        for (int k = 0; k < kStage; k++)
            for (cudaEvent_t *ev : { &groups[g].in_free[k], &groups[g].out_free[k] }) if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
    }
    if (fork_ev) { cudaEventDestroy(fork_ev); fork_ev = nullptr; }
    if (out_done) { cudaEventDestroy(out_done); out_done = nullptr; }
    for (cudaStream_t *st : { &sp_fwd, &sp_inv }) if (*st) { cudaStreamSynchronize(*st); cudaStreamDestroy(*st); *st = nullptr; }
    for (int k = 0; k < 4; k++) for (cudaEvent_t *ev : { &q_in_ready[k], &q_out_ready[k] }) if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
    for (int k = 0; k < 2; k++) {
        for (cudaEvent_t *ev : { &sp_fwd_done[k], &sp_mac_done[k], &sp_inv_done[k], &sp_arrived[k] }) if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
        for (int j = 0; j < 8; j++) if (sp_acc[k][j]) { cudaFree(sp_acc[k][j]); sp_acc[k][j] = nullptr; }
    }
    if (tail_stream) { cudaStreamSynchronize(tail_stream); cudaStreamDestroy(tail_stream); tail_stream = nullptr; }
    for (int g = 0; g < BFIR_MAX_GROUPS; g++) if (tail_done[g]) { cudaEventDestroy(tail_done[g]); tail_done[g] = nullptr; }
    if (tail_join_ev) { cudaEventDestroy(tail_join_ev); tail_join_ev = nullptr; }
    for (int k = 0; k < kMaxInflight; k++)
        for (int g = 0; g < BFIR_MAX_GROUPS; g++) if (ticket_ev[k][g]) { cudaEventDestroy(ticket_ev[k][g]); ticket_ev[k][g] = nullptr; }
    if (stream && own_stream) cudaStreamDestroy(stream);
    stream = nullptr;
    for (int q = 0; q < BFIR_MAX_PEERS; q++) if (peer_opened[q]) { cudaIpcCloseMemHandle(peer_opened[q]); peer_opened[q] = nullptr; }
    if (recv) { cudaFree(recv); recv = nullptr; }
    for (int k = 1; k < kStage; k++) { if (stage_in[k]) cudaFree(stage_in[k]); if (stage_out[k]) cudaFree(stage_out[k]); stage_in[k] = stage_out[k] = nullptr; }
    stage_in[0] = stage_out[0] = nullptr;
    void *bufs[] = { plan8, xin8, out8, acc_oct[0], acc_oct[1], acc_oct[2], acc_oct[3], d_peer_timeout, coeff_map, acc_quad[0], acc_quad[1], acc_pair, fdl, coeffs, acc, prev, ybuf, tw, d_in, d_out, state, procblocks, coeff_blocks, pb_inc, nonfinite, stats, xin, yacc, gains_in, gains_out, coeffs_next, acc2, tbuf };
    for (void *b : bufs) if (b) cudaFree(b);
    out8 = plan8 = xin8 = nullptr;
    acc_pair = nullptr; acc_quad[0] = acc_quad[1] = nullptr; acc_oct[0] = acc_oct[1] = acc_oct[2] = acc_oct[3] = nullptr;
    fdl = coeffs = acc = prev = ybuf = tw = d_in = d_out = xin = yacc = gains_in = gains_out = coeffs_next = acc2 = tbuf = nullptr;
    state = nullptr; procblocks = coeff_blocks = nonfinite = nullptr; coeff_map = nullptr; d_peer_timeout = nullptr; pb_inc = nullptr; stats = nullptr;
    if (h_state) cudaFreeHost(h_state);
    h_state = nullptr;
    if (h_flag) cudaFreeHost(h_flag);
    h_flag = nullptr; d_flag = nullptr;
    dither.destroy();
}

int Engine::reset()
{
    // brutefir::reset, brutefir.cpp:347-367
    const int nmax = Cot > Ct ? Cot : Ct;
    const int threads = 256, blocks = (nmax + threads - 1) / threads;
    engine_reset_kernel<<<blocks, threads, 0, stream>>>(state, procblocks, stats, Ct, nmax);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    for (auto &o : last_overflow) { o.n_overflows = 0; o.largest = 0; o.intlargest = 0; }
    host_blockcounter = 0;
    tail_ready = false;
    return BFIR_OK;
}

int Engine::set_coeff(const void *const *h_coeffs, int n_coeffs, int length, int blocks, double scale)
{
    return load_coeff(h_coeffs, nullptr, 0, n_coeffs, length, blocks, scale, false);
}

// brutefir::set_coeff, brutefir.cpp:180-228 -> coeff::preprocess_coeff, coeff.cpp:293-354
//   -> convolver_coeffs2cbuf, fftw_convolver.cpp:475-537, for every (channel, partition) in ONE launch.
// Source: planar HOST arrays (h_coeffs[n]) or one planar DEVICE array (d_src + n * d_stride elements;
// stride 0 = the same filter for every channel, as the equalizer produces). next = true stages the set
// for a cross-faded swap on the following block instead of replacing the current one.
int Engine::load_coeff(const void *const *h_coeffs, const void *d_src, long long d_stride, int n_coeffs, int length, int blocks, double scale, bool next)
{
    if ((h_coeffs == nullptr && d_src == nullptr) || blocks < 1 || length < 0) { set_error("invalid coefficient arguments"); return BFIR_ERR_INVALID; }
    BFIR_CUDA(cudaSetDevice(device));
    BFIR_CUDA(cudaStreamSynchronize(stream));
    invalidate_graphs();
    tail_ready = false;
    if (n_coeffs > Ct) n_coeffs = Ct;                     // brutefir.cpp:191-194
    const size_t cbuf = (size_t)N * rs;
    void *target = nullptr;
    if (next) {
        if (!initialized || blocks != coeff_alloc || n_coeffs != Ct) { set_error("crossfade swap needs an initialised engine and the same filter geometry"); return BFIR_ERR_INVALID; }
        if (xbar || part_count != P) { set_error("crossfade swap is not available with a crossbar or a partition shard"); return BFIR_ERR_INVALID; }
        if (!coeffs_next) BFIR_CUDA(cudaMalloc(&coeffs_next, cbuf * blocks * Ct));
        if (!acc2) BFIR_CUDA(cudaMalloc(&acc2, cbuf * Ct));
        if (!tbuf) BFIR_CUDA(cudaMalloc(&tbuf, 2 * cbuf * Ct));
        target = coeffs_next;
    } else {
        initialized = false;                               // free_coeff(), brutefir.cpp:189
        xfade_pending = false;
        if (coeffs) { cudaFree(coeffs); coeffs = nullptr; }
        if (coeffs_next) { cudaFree(coeffs_next); coeffs_next = nullptr; }
        BFIR_CUDA(cudaMemsetAsync(coeff_blocks, 0, sizeof(int) * Ct, stream));
        if (n_coeffs < 1) return BFIR_OK;
        coeff_alloc = blocks;
        BFIR_CUDA(cudaMalloc(&coeffs, cbuf * blocks * Ct));
        target = coeffs;
    }
    BFIR_CUDA(cudaMemsetAsync(target, 0, cbuf * blocks * Ct, stream));
    void *d_planar = nullptr;
    long long stride = d_stride;
    if (d_src == nullptr) {                                // stage the planar host arrays
        const size_t row = (size_t)(length > 0 ? length : 1) * rs;
        BFIR_CUDA(cudaMalloc(&d_planar, row * n_coeffs));
        for (int n = 0; n < n_coeffs; n++) {
            if (h_coeffs[n] == nullptr) { cudaFree(d_planar); set_error("coefficient array %d is NULL", n); return BFIR_ERR_INVALID; }
            if (length > 0) BFIR_CUDA(cudaMemcpyAsync((char *)d_planar + row * n, h_coeffs[n], (size_t)length * rs, cudaMemcpyHostToDevice, stream));
        }
        stride = (long long)(length > 0 ? length : 1);
    }
    BFIR_CUDA(cudaMemsetAsync(nonfinite, 0, sizeof(int), stream));
    FwdArgs a = {};
    a.in_mode = IN_COEFF; a.out_layout = LAYOUT_ORD;
    a.in = d_src ? d_src : d_planar; a.in_stride_x = stride; a.in_stride_y = 0;
    a.out = target; a.out_stride_x = (long long)blocks * N; a.out_stride_y = N;
    a.scale_in = scale; a.scale_out = 1.0 / (double)N;     // fftw_convolver.cpp:520
    a.coeff_len = length; a.nonfinite = nonfinite;
    cudaError_t e = launch_rfft_forward(rs, log2m, rfft_choose_r0(rs, log2m, (long long)n_coeffs * blocks), dim3(n_coeffs, blocks), stream, a, tw);
    count_launch();
    if (e != cudaSuccess) { if (d_planar) cudaFree(d_planar); set_error("coefficient transform launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    int bad = 0;
    if (!next) {
        std::vector<int> hb(Ct, 0);
        for (int n = 0; n < n_coeffs; n++) hb[n] = blocks; // brutefir.cpp:213
        BFIR_CUDA(cudaMemcpyAsync(coeff_blocks, hb.data(), sizeof(int) * Ct, cudaMemcpyHostToDevice, stream));
    }
    BFIR_CUDA(cudaMemcpyAsync(&bad, nonfinite, sizeof(int), cudaMemcpyDeviceToHost, stream));
    BFIR_CUDA(cudaStreamSynchronize(stream));
    if (d_planar) cudaFree(d_planar);
    if (bad) {                                             // brutefir.cpp:207-224
        pinfo("NaN or Inf value among coefficients.\n");
        set_error("NaN or Inf value among coefficients.");
        if (!next) { cudaFree(coeffs); coeffs = nullptr; cudaMemset(coeff_blocks, 0, sizeof(int) * Ct); }
        return BFIR_ERR_COEFF;
    }
    if (next) xfade_pending = true;
    else initialized = true;                               // brutefir.cpp:226
    return BFIR_OK;
}

// after a block has been enqueued: a staged coefficient set becomes the current one
void Engine::finish_block()
{
    if (xfade_pending) { void *t = coeffs; coeffs = coeffs_next; coeffs_next = t; xfade_pending = false; invalidate_graphs(); }
    blocks_since_sync++;
    host_blockcounter++;
}

int Engine::set_crossbar(const double *in_gains, const double *out_gains)
{
    if (!xbar) { set_error("engine was created without a crossbar"); return BFIR_ERR_INVALID; }
    if (in_gains == nullptr || out_gains == nullptr) return BFIR_ERR_INVALID;
    BFIR_CUDA(cudaStreamSynchronize(stream));
    for (int g = 0; g < n_groups; g++) BFIR_CUDA(cudaStreamSynchronize(groups[g].stream));
    // like mixnscalef (fftw_convolver.cpp:871-876) the float engine narrows the gains to float first
    std::vector<unsigned char> a((size_t)C * Ci * rs), b((size_t)Co * C * rs);
    for (int k = 0; k < C * Ci; k++) { if (rs == 4) ((float *)a.data())[k] = (float)in_gains[k]; else ((double *)a.data())[k] = in_gains[k]; }
    for (int k = 0; k < Co * C; k++) { if (rs == 4) ((float *)b.data())[k] = (float)out_gains[k]; else ((double *)b.data())[k] = out_gains[k]; }
    BFIR_CUDA(cudaMemcpy(gains_in, a.data(), a.size(), cudaMemcpyHostToDevice));
    BFIR_CUDA(cudaMemcpy(gains_out, b.data(), b.size(), cudaMemcpyHostToDevice));
    xbar_set = true;
    return BFIR_OK;
}

// fused reduce set-up: reduced channels (outputs with a crossbar, else filters) are dealt to the ranks
// in contiguous ranges of cpr = ceil(n / world); allocates this rank's receive buffer
int Engine::peer_setup(int rank, int world)
{
    const int n_red = xbar ? Cot : Ct;
    if (world < 2 || world > BFIR_MAX_PEERS || rank < 0 || rank >= world) { set_error("invalid peer geometry"); return BFIR_ERR_INVALID; }
    if (S != 1 || dither_on || world > n_red) { set_error("fused reduce needs one stream, no dither and world <= channels"); return BFIR_ERR_INVALID; }
    BFIR_CUDA(cudaStreamSynchronize(stream));
    memset(&peer, 0, sizeof(peer));
    peer.world = world; peer.self = rank; peer.cpr = (n_red + world - 1) / world;
    own_first = rank * peer.cpr;
    own_count = own_first >= n_red ? 0 : (n_red - own_first < peer.cpr ? n_red - own_first : peer.cpr);
    if (own_count < 1) { set_error("rank %d owns no channel", rank); return BFIR_ERR_INVALID; }
    if (recv) { cudaFree(recv); recv = nullptr; }
    const size_t data_bytes = (size_t)BFIR_PEER_PHASES * world * peer.cpr * N * rs;
    size_t bytes = data_bytes + 256;                       // + arrival flags (outputs at +0, inputs at +128)
    peer.flag_offset = (long long)data_bytes;
    peer.flag_in_offset = (long long)data_bytes + 128;
    if (const char *env = getenv("BFIR_SHARD_INPUTS")) shard_inputs = atoi(env) != 0;
    peer.xin_offset = 0; peer.n_inputs = Cit; peer.in_cpr = (Cit + world - 1) / world;
    own_in_first = rank * peer.in_cpr;
    own_in_count = own_in_first >= Cit ? 0 : (Cit - own_in_first < peer.in_cpr ? Cit - own_in_first : peer.in_cpr);
    if (xbar && shard_inputs && S == 1) {                  // input region: [2 call parities * 8 blocks][Cit][N]
        peer.xin_offset = (long long)bytes;
        bytes += (size_t)16 * Cit * N * rs;
    }
    BFIR_CUDA(cudaMalloc(&recv, bytes));
    BFIR_CUDA(cudaMemset(recv, 0, bytes));
    if (!d_peer_timeout) { BFIR_CUDA(cudaMalloc((void **)&d_peer_timeout, sizeof(int))); BFIR_CUDA(cudaMemset(d_peer_timeout, 0, sizeof(int))); }
    peer_epoch = 0; peer_in_epoch = 0; peer_quad_pending = false;
    peer.recv[rank] = recv;
    peer.enabled = 1;
    invalidate_graphs();
    set_groups(1);
    return BFIR_OK;
}

int Engine::set_groups(int n)
{
    if (n < 1) n = 1;
    if (n > BFIR_MAX_GROUPS) n = BFIR_MAX_GROUPS;
    if (n > S) n = S;
    if (stream) BFIR_CUDA(cudaStreamSynchronize(stream));
    for (int g = 0; g < n_groups && g < BFIR_MAX_GROUPS; g++)
        if (groups[g].stream) BFIR_CUDA(cudaStreamSynchronize(groups[g].stream));
    n_groups = n;
    done_ticket = next_ticket;      // everything queued so far has completed (streams synchronised above)
    tail_ready = false;
    invalidate_graphs();
    const int base = S / n, extra = S % n;
    int s = 0;
    for (int g = 0; g < n; g++) {
        const int cnt = base + (g < extra ? 1 : 0);
        groups[g].s0 = s; groups[g].s1 = s + cnt;
        groups[g].c0 = s * C; groups[g].c1 = (s + cnt) * C;
        s += cnt;
    }
    // all group counters continue from group 0's value
    EngineState st;
    BFIR_CUDA(cudaMemcpy(&st, state, sizeof(st), cudaMemcpyDeviceToHost));
    EngineState all[BFIR_MAX_GROUPS];
    for (int g = 0; g < BFIR_MAX_GROUPS; g++) all[g] = st;
    BFIR_CUDA(cudaMemcpy(state, all, sizeof(all), cudaMemcpyHostToDevice));
    return BFIR_OK;
}

// group streams start after everything queued on the engine's stream ...
int Engine::fork()
{
    if (n_groups == 1) return BFIR_OK;
    BFIR_CUDA(cudaEventRecord(fork_ev, stream));
    for (int g = 0; g < n_groups; g++) BFIR_CUDA(cudaStreamWaitEvent(groups[g].stream, fork_ev, 0));
    return BFIR_OK;
}

// ... and the engine's stream continues after every group is done
int Engine::join()
{
    if (n_groups == 1) return BFIR_OK;
    for (int g = 0; g < n_groups; g++) {
        BFIR_CUDA(cudaEventRecord(groups[g].done, groups[g].stream));
        BFIR_CUDA(cudaStreamWaitEvent(stream, groups[g].done, 0));
    }
    return BFIR_OK;
}

// input FFT into the delay line + this engine's partition sum, for the channels of one group
int Engine::front_group(int g, const void *d_inbuf, cudaEvent_t *input_consumed, bool skip_mac)
{
    const Group &grp = groups[g];
    const int s0 = n_groups == 1 ? 0 : grp.s0, ns = n_groups == 1 ? S : grp.s1 - grp.s0;
    const int nch = ns * C, c0 = s0 * C;
    cudaStream_t st = stage_stream ? stage_stream : gstream(g);
    FwdArgs f = {};
    f.in_mode = IN_RAW_PREV; f.out_layout = LAYOUT_ORD;
    f.in = d_inbuf; f.in_stride_x = (long long)L * Ci * in_sf.bytes;  // bytes per stream
    f.scale_in = 1.0; f.scale_out = in_sf.scale;                       // brutefir.cpp:273-277
    f.prev = prev; f.fmt = in_sf.format; f.ch_per_stream = Ci; f.n_channels = Cit; f.ch_base = s0 * Ci;
    f.use_abs_block = use_abs ? 1 : 0; f.abs_block = host_blockcounter + (unsigned int)fwd_block_offset;
    f.state = state + g; f.n_slots = Pslots; f.n_parts = P; f.prev_parity = (host_blockcounter + (unsigned int)fwd_block_offset) & 1u; f.slot_offset = fwd_block_offset;
    if (!xbar) { f.out = fdl; f.out_stride_x = (long long)Pslots * N; f.out_stride_y = N; f.procblocks = procblocks; f.pb_inc = pb_inc; }
    else { f.out = xin; f.out_stride_x = N; f.out_stride_y = 0; f.procblocks = nullptr; f.pb_inc = nullptr; }
    if (g == 0 && !prof_suppress) prof_nb = 1;
    if (g == 0) prof(0);
    cudaError_t e = launch_rfft_forward(rs, log2m, fft_r0, dim3(ns * Ci, 1), st, f, tw);
    count_launch();
    if (e != cudaSuccess) { set_error("forward launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    if (input_consumed) BFIR_CUDA(cudaEventRecord(*input_consumed, st));   // the raw input block has been read
    if (xbar) { // inputs -> filter inputs, straight into the delay-line slot (mixnscale INPUT, n_bufs = Ci)
        XbarArgs x = {};
        x.in = xin; x.in_stride = N; x.out = fdl; x.out_stride = (long long)Pslots * N; x.slot_stride = N;
        x.gains = gains_in; x.n_in = Ci; x.n_out = C; x.N = N; x.n_streams = ns; x.stream_base = s0;
        x.state = state + g; x.n_slots = Pslots; x.n_parts = P; x.slot_offset = fwd_block_offset; x.procblocks = procblocks; x.pb_inc = pb_inc;
        x.use_abs_block = use_abs ? 1 : 0; x.abs_block = host_blockcounter + (unsigned int)fwd_block_offset;
        xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(Ci) : xbar_kernel_for<double>(Ci);
        xk<<<dim3((N + 255) / 256, ns), 256, (size_t)C * Ci * rs, st>>>(x);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    }
    if (g == 0) prof(1);
    if (skip_mac) return BFIR_OK;   // the look-ahead launch has left sum_{i>=1} in acc; back_group adds the head term
    tail_ready = false;             // a full partition sum overwrites whatever a look-ahead launch left in acc

    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs; m.acc = acc;
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = part_begin; m.part_count = part_count;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state + g; m.block_offset = 0; m.ch_base = c0;
    if (peer.enabled && !xbar) m.push = peer;
    dim3 grid((N / 8 + 256 / mac_split - 1) / (256 / mac_split), nch);
    mac_kernel_t mk = rs == 4 ? mac_kernel_for_split<float>(mac_split) : mac_kernel_for_split<double>(mac_split);
    mk<<<grid, 256, 0, st>>>(m);
    if (xfade_pending) {                                                // same block through the staged filters
        m.coeffs = coeffs_next; m.acc = acc2;
        mk<<<grid, 256, 0, st>>>(m);
        count_launch();
    }
    if (g == 0) prof(2);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    if (peer.enabled && xbar) { // partial out-mix on this rank's partition shard, rows pushed to their owners
        XbarArgs x = {};
        x.in = acc; x.in_stride = N; x.out = yacc; x.out_stride = N; x.slot_stride = 0;
        x.gains = gains_out; x.n_in = C; x.n_out = Co; x.N = N; x.n_streams = ns; x.stream_base = s0;
        x.state = nullptr; x.n_slots = Pslots; x.n_parts = P; x.push = peer; x.push_state = state + g; x.push_phase = -1;
        xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(C) : xbar_kernel_for<double>(C);
        xk<<<dim3((N + 255) / 256, ns), 256, (size_t)Co * C * rs, st>>>(x);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    }
    return BFIR_OK;
}

// output stage from the accumulated spectra, for the channels of one group
int Engine::back_group(int g, void *d_outbuf, bool head)
{
    const Group &grp = groups[g];
    const int s0 = n_groups == 1 ? 0 : grp.s0, ns = n_groups == 1 ? S : grp.s1 - grp.s0;
    const int nch = ns * Co, c0 = s0 * Co;
    cudaStream_t st = stage_stream ? stage_stream : gstream(g);
    if (peer.enabled) { // owner side: sum the source slots of the own channels, then the normal output stage on them
        void *dst = xbar ? yacc : acc;
        dim3 grid((N + 255) / 256, own_count);
        if (rs == 4) peer_sum_kernel<float><<<grid, 256, 0, st>>>((const float *)recv, (float *)dst, state + g, peer.world, peer.cpr, N, own_first, own_count, peer_phase);
        else peer_sum_kernel<double><<<grid, 256, 0, st>>>((const double *)recv, (double *)dst, state + g, peer.world, peer.cpr, N, own_first, own_count, peer_phase);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
        InvArgs v = {};
        v.in_layout = LAYOUT_ORD; v.in = dst; v.in_stride_x = N; v.scale_in = out_sf.scale;
        v.fmt = out_sf.format; v.ch_per_stream = own_count; v.ovf_max = ovf_max; v.stats = stats; v.state = state + g; v.host_flag = d_flag;
        v.ch_base = own_first; v.raw_ch_base = own_first;
        v.out_mode = OUT_RAW; v.out = d_outbuf; v.out_stride_x = (long long)L * own_count * out_sf.bytes;
        cudaError_t e = launch_rfft_inverse(rs, log2m, fft_r0, dim3(own_count, 1), st, v, tw);
        count_launch();
        if (e != cudaSuccess) { set_error("inverse launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
        if (g == 0) { prof(3); prof_next(); }
        return BFIR_OK;
    }
    if (xbar) { // filter outputs -> outputs (mixnscale OUTPUT, n_bufs = C)
        XbarArgs x = {};
        x.in = acc_override ? acc_override : acc; x.in_stride = N; x.out = yacc; x.out_stride = N; x.slot_stride = 0;
        x.gains = gains_out; x.n_in = C; x.n_out = Co; x.N = N; x.n_streams = ns; x.stream_base = s0;
        x.state = nullptr; x.n_slots = Pslots; x.n_parts = P; x.procblocks = nullptr; x.pb_inc = nullptr;
        xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(C) : xbar_kernel_for<double>(C);
        xk<<<dim3((N + 255) / 256, ns), 256, (size_t)Co * C * rs, st>>>(x);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    }
    if (xfade_pending) { // old and new filter outputs -> time domain -> linear ramp fused with the output stage
        InvArgs t = {};
        t.in_layout = LAYOUT_ORD; t.in_stride_x = N; t.scale_in = out_sf.scale; t.out_mode = OUT_TIME; t.out_stride_x = N; t.ch_base = c0;
        for (int k = 0; k < 2; k++) {
            t.in = k == 0 ? acc : acc2;
            t.out = (char *)tbuf + (size_t)k * N * rs * Ct;
            cudaError_t e = launch_rfft_inverse(rs, log2m, fft_r0, dim3(nch, 1), st, t, tw);
            count_launch();
            if (e != cudaSuccess) { set_error("inverse launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
        }
        XfadeArgs x = {};
        x.t_old = tbuf; x.t_new = (char *)tbuf + (size_t)N * rs * Ct;
        x.out = dither_on ? ybuf : d_outbuf; x.out_stream_stride = (long long)L * Co * out_sf.bytes;
        x.N = N; x.L = L; x.fmt = out_sf.format; x.ch_per_stream = Co; x.ch_base = c0; x.to_real = dither_on ? 1 : 0;
        x.ovf_max = ovf_max; x.stats = stats; x.state = state + g; x.host_flag = d_flag;
        if (rs == 4) xfade_emit_kernel<float><<<dim3((L + 255) / 256, nch), 256, 0, st>>>(x);
        else xfade_emit_kernel<double><<<dim3((L + 255) / 256, nch), 256, 0, st>>>(x);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    }
    InvArgs v = {};
    v.in_layout = LAYOUT_ORD; v.in = xbar ? yacc : (acc_override ? acc_override : acc); v.in_stride_x = N;
    v.scale_in = out_sf.scale;                                         // brutefir.cpp:303-307
    v.fmt = out_sf.format; v.ch_per_stream = Co; v.ovf_max = ovf_max; v.stats = stats; v.state = state + g; v.ch_base = c0; v.host_flag = d_flag;
    if (dither_on) { v.out_mode = OUT_REAL_L; v.out = ybuf; v.out_stride_x = L; }
    else { v.out_mode = OUT_RAW; v.out = d_outbuf; v.out_stride_x = (long long)L * Co * out_sf.bytes; }
    if (head) {
        v.head_x = fdl; v.head_x_stride = (long long)Pslots * N;
        v.head_h = coeffs; v.head_h_stride = (long long)coeff_alloc * N;
        v.head_blocks = coeff_blocks; v.head_map = coeff_map;
    }
    if (!xfade_pending) {
        cudaError_t e = launch_rfft_inverse(rs, log2m, fft_r0, dim3(nch, 1), st, v, tw);
        count_launch();
        if (e != cudaSuccess) { set_error("inverse launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    }
    if (dither_on) {
        DitherArgs d = {};
        d.real = ybuf; d.real_stride = L; d.raw = d_outbuf; d.raw_stream_stride = (long long)L * Co * out_sf.bytes;
        d.fmt = out_sf.format; d.ch_per_stream = Co; d.L = L; d.n_channels = nch; d.ch_base = c0;
        d.randtab = dither.d_tab; d.randtab_size = dither.size; d.randmap = dither.d_map;
        d.dstate = dither.d_state; d.stats = stats; d.single_channel = -1;
        const int G = dither_channels_per_cta(nch), blocks = (nch + G - 1) / G;
        if (rs == 4) dither_kernel<float><<<blocks, 128, 0, st>>>(d, G);
        else dither_kernel<double><<<blocks, 128, 0, st>>>(d, G);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    }
    if (g == 0 && !prof_suppress) { prof(3); prof_next(); }
    return BFIR_OK;
}

// Input stage of nb consecutive blocks (one group, stage pipeline, host-given block index) of a CROSSBAR engine in three
// launches: the raw blocks de-interleaved into planar rows (which makes the blocks independent of each other: block b's
// previous block is plan[b-1], block 0's the engine's previous-block row), ONE forward-transform launch over all inputs of
// all blocks (grid.y = block), ONE input-mix launch (grid.z = block) into the delay-line slots. Without a crossbar (128
// transforms per block already fill the GPU, and one launch per block leaves SMs to the partition sum running beside them)
// and for fewer than two blocks: block by block.
int Engine::front_blocks(int nb, const void *const *d_in, cudaStream_t st)
{
    if (!batch_stages || !xbar || nb < 2 || nb > 8 || S != 1) {
        int rc = BFIR_OK;
        cudaStream_t keep = stage_stream;
        stage_stream = st;
        for (int b = 0; b < nb && rc == BFIR_OK; b++) { fwd_block_offset = b; rc = front_group(0, d_in[b], nullptr, true); }
        fwd_block_offset = 0;
        stage_stream = keep;
        return rc;
    }
    if (!plan8) BFIR_CUDA(cudaMalloc(&plan8, (size_t)8 * Cit * L * rs));
    if (!xin8) BFIR_CUDA(cudaMalloc(&xin8, (size_t)8 * Cit * N * rs));
    const unsigned int t0 = host_blockcounter;
    PlanarArgs pa = {};
    for (int b = 0; b < nb; b++) pa.raw[b] = d_in[b];
    pa.plan = plan8; pa.L = L; pa.n_ch = Cit; pa.fmt = in_sf.format; pa.nb = nb;
    if (rs == 4) raw_to_planar_kernel<float><<<dim3((L + 255) / 256, Cit, nb), 256, 0, st>>>(pa);
    else raw_to_planar_kernel<double><<<dim3((L + 255) / 256, Cit, nb), 256, 0, st>>>(pa);
    count_launch();
    FwdArgs f = {};
    f.in_mode = IN_PLANAR2; f.out_layout = LAYOUT_ORD;
    f.scale_in = 1.0; f.scale_out = in_sf.scale;
    f.out = xin8; f.out_stride_x = N; f.out_stride_y = (long long)Cit * N;
    f.n_channels = Cit; f.ch_per_stream = Ci; f.fmt = in_sf.format;
    for (int b = 0; b < nb; b++) {
        f.hi_multi[b] = (char *)plan8 + (size_t)b * Cit * L * rs;
        f.lo_multi[b] = b == 0 ? (const void *)((char *)prev + ((size_t)(t0 & 1u) * Cit) * L * rs) : (const void *)((char *)plan8 + (size_t)(b - 1) * Cit * L * rs);
    }
    cudaError_t e = launch_rfft_forward(rs, log2m, fft_r0, dim3(Cit, nb), st, f, tw);
    count_launch();
    if (e != cudaSuccess) { set_error("forward launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    // the last block of the call goes where every other entry point expects the previous block -- AFTER the transforms: for an
    // even number of blocks that is the very row block 0 has just read as ITS previous block
    BFIR_CUDA(cudaMemcpyAsync((char *)prev + ((size_t)(((t0 + (unsigned int)nb - 1u) & 1u) ^ 1u) * Cit) * L * rs,
                              (char *)plan8 + (size_t)(nb - 1) * Cit * L * rs, (size_t)Cit * L * rs, cudaMemcpyDeviceToDevice, st));
    XbarArgs x = {};
    x.in_stride = N; x.out = fdl; x.out_stride = (long long)Pslots * N; x.slot_stride = N;
    x.gains = gains_in; x.n_in = Ci; x.n_out = C; x.N = N; x.n_streams = S; x.stream_base = 0;
    x.state = state; x.n_slots = Pslots; x.n_parts = P; x.slot_offset = 0; x.procblocks = procblocks; x.pb_inc = pb_inc;
    x.use_abs_block = 1; x.abs_block = t0;
    x.n_multi = nb;
    for (int b = 0; b < nb; b++) { x.in_multi[b] = (char *)xin8 + (size_t)b * Cit * N * rs; x.out_multi[b] = nullptr; }
    xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(Ci) : xbar_kernel_for<double>(Ci);
    xk<<<dim3((N + 255) / 256, S, nb), 256, (size_t)C * Ci * rs, st>>>(x);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    return BFIR_OK;
}

// Output stage of nb consecutive blocks (one group, stage pipeline) in as few launches as possible: on a partition shard ONE
// slot-sum launch (grid.z = block), with a crossbar ONE output-mix launch, then ONE inverse-transform launch (grid.y =
// block) -- instead of nb times (sum / mix + transform) of latency-bound kernels with at most Co CTAs each. acc_in[b]:
// accumulated filter-output spectra of block b (ignored on a shard, where the receive-buffer phases peer_phase + b are
// summed). Dither and a pending filter swap take the block-by-block path.
int Engine::back_blocks(int nb, void *const *acc_in, void *const *d_out, cudaStream_t st)
{
    // (without a crossbar or shard a block's transforms nearly fill the GPU; one launch per block then leaves the remaining SMs
    //  to the partition sum running beside them, and batching measured 3 % slower on cfg1 x 16)
    if (!batch_stages || dither_on || xfade_pending || nb < 2 || nb > 8 || !(xbar || peer.enabled)) {
        int rc = BFIR_OK;
        cudaStream_t keep = stage_stream;
        stage_stream = st;
        const int base = peer_phase;
        for (int b = 0; b < nb && rc == BFIR_OK; b++) {
            if (peer.enabled) peer_phase = base + b; else acc_override = acc_in[b];
            rc = back_group(0, d_out[b]);
        }
        acc_override = nullptr;
        peer_phase = base;
        stage_stream = keep;
        return rc;
    }
    const size_t cbuf = (size_t)N * rs;
    const int nred = xbar ? Cot : Ct;
    if (xbar || peer.enabled) {
        const size_t need = cbuf * (size_t)(Cot > Ct ? Cot : Ct);
        if (!out8) { BFIR_CUDA(cudaMalloc(&out8, need * 8)); out8_stride = need; }
    }
    InvArgs v = {};
    v.in_layout = LAYOUT_ORD; v.in_stride_x = N; v.scale_in = out_sf.scale;
    v.fmt = out_sf.format; v.ovf_max = ovf_max; v.stats = stats; v.state = state; v.host_flag = d_flag;
    v.out_mode = OUT_RAW; v.n_multi = nb;
    int n_launch_ch;
    if (peer.enabled) {   // owner side: sum the source slots of the own channels of all nb blocks
        dim3 grid((N + 255) / 256, own_count, nb);
        const long long bstride = (long long)(out8_stride / rs);
        if (rs == 4) peer_sum_kernel<float><<<grid, 256, 0, st>>>((const float *)recv, (float *)out8, state, peer.world, peer.cpr, N, own_first, own_count, peer_phase, bstride);
        else peer_sum_kernel<double><<<grid, 256, 0, st>>>((const double *)recv, (double *)out8, state, peer.world, peer.cpr, N, own_first, own_count, peer_phase, bstride);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
        for (int b = 0; b < nb; b++) { v.in_multi[b] = (char *)out8 + (size_t)b * out8_stride; v.out_multi[b] = d_out[b]; }
        v.ch_per_stream = own_count; v.ch_base = own_first; v.raw_ch_base = own_first;
        v.out_stride_x = (long long)L * own_count * out_sf.bytes;
        n_launch_ch = own_count;
    } else {
        if (xbar) {       // filter outputs -> outputs of all nb blocks (mixnscale OUTPUT, n_bufs = C)
            XbarArgs x = {};
            x.in_stride = N; x.out = out8; x.out_stride = N; x.slot_stride = 0;
            x.gains = gains_out; x.n_in = C; x.n_out = Co; x.N = N; x.n_streams = S; x.stream_base = 0;
            x.state = nullptr; x.n_slots = Pslots; x.n_parts = P; x.procblocks = nullptr; x.pb_inc = nullptr;
            x.n_multi = nb;
            for (int b = 0; b < nb; b++) { x.in_multi[b] = acc_in[b]; x.out_multi[b] = (char *)out8 + (size_t)b * out8_stride; }
            xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(C) : xbar_kernel_for<double>(C);
            xk<<<dim3((N + 255) / 256, S, nb), 256, (size_t)Co * C * rs, st>>>(x);
            count_launch();
            BFIR_CUDA(cudaGetLastError());
        }
        for (int b = 0; b < nb; b++) { v.in_multi[b] = xbar ? (void *)((char *)out8 + (size_t)b * out8_stride) : acc_in[b]; v.out_multi[b] = d_out[b]; }
        v.ch_per_stream = Co; v.ch_base = 0; v.raw_ch_base = 0;
        v.out_stride_x = (long long)L * Co * out_sf.bytes;
        n_launch_ch = S * Co;
    }
    (void)nred;
    cudaError_t e = launch_rfft_inverse(rs, log2m, fft_r0, dim3(n_launch_ch, 1), st, v, tw);
    count_launch();
    if (e != cudaSuccess) { set_error("inverse launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    return BFIR_OK;
}

// look-ahead: partitions 1 .. P-1 of the NEXT block for the channels of one group (blockcounter has been advanced
// by the inverse kernel of the block that just went out; the next forward transform has not counted itself yet)
int Engine::tail_group(int g, cudaStream_t st)
{
    const Group &grp = groups[g];
    const int s0 = n_groups == 1 ? 0 : grp.s0, ns = n_groups == 1 ? S : grp.s1 - grp.s0;
    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs; m.acc = acc;
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = 1; m.part_count = P - 1;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state + g; m.block_offset = 0; m.ch_base = s0 * C;
    m.procblocks_bias = 1;
    dim3 grid((N / 8 + 256 / mac_split - 1) / (256 / mac_split), ns * C);
    mac_kernel_t mk = rs == 4 ? mac_kernel_for_split<float>(mac_split) : mac_kernel_for_split<double>(mac_split);
    mk<<<grid, 256, 0, st>>>(m);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    return BFIR_OK;
}

// both blocks of a pair for the channels of one group: forward t, forward t+1, ONE partition sum for both,
// inverse t, inverse t+1 (each inverse advances the block counter)
int Engine::pair_group(int g, const void *d_in0, const void *d_in1, void *d_out0, void *d_out1, cudaEvent_t *input_consumed, cudaEvent_t *output_free)
{
    const Group &grp = groups[g];
    const int s0 = n_groups == 1 ? 0 : grp.s0, ns = n_groups == 1 ? S : grp.s1 - grp.s0;
    // profiling (bfir_set_profiling): one entry per PAIR -- both forward transforms, the pair sum, both inverse transforms
    if (g == 0) prof_nb = 2;
    if (g == 0) prof(0);
    prof_suppress = true;
    int rc = front_group(g, d_in0, nullptr, true);
    fwd_block_offset = 1;
    if (rc == BFIR_OK) rc = front_group(g, d_in1, input_consumed, true);
    fwd_block_offset = 0;
    prof_suppress = false;
    if (rc != BFIR_OK) return rc;
    if (g == 0) prof(1);
    tail_ready = false;
    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs; m.acc = acc; m.acc_next = acc_pair;
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = 0; m.part_count = P;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state + g; m.block_offset = 0; m.ch_base = s0 * C;
    dim3 grid((N / 8 + 256 / mac_split - 1) / (256 / mac_split), ns * C);
    mac_kernel_t mk = rs == 4 ? mac_pair_kernel_for_split<float>(mac_split) : mac_pair_kernel_for_split<double>(mac_split);
    mk<<<grid, 256, 0, gstream(g)>>>(m);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    if (g == 0) prof(2);
    if (output_free) BFIR_CUDA(cudaStreamWaitEvent(gstream(g), *output_free, 0));   // the staging slices have been copied out
    prof_suppress = true;
    rc = back_group(g, d_out0);
    acc_override = acc_pair;
    if (rc == BFIR_OK) rc = back_group(g, d_out1);
    acc_override = nullptr;
    prof_suppress = false;
    if (g == 0) { prof(3); prof_next(); }
    return rc;
}

// One call of NB = 2 or 4 consecutive blocks through the stage pipeline (one group). Call k = blocks t .. t+NB-1,
// t = host_blockcounter:
//   forward stream:  after the partition sum of call k-2 (which still reads the slots these transforms overwrite;
//                    call k-1's sum reaches back to block t-P-3 at most, and the delay line has P+7 slots, so the
//                    slots of blocks t .. t+3 are not among its operands): forward t .. t+NB-1
//   engine's stream: after those transforms and after the inverse transforms of call k-2 (same accumulators): ONE
//                    partition-sum launch for the NB blocks (pair / multi kernel)
//   inverse stream:  after the sum: inverse t .. t+NB-1 (these advance the device block counter)
// The inputs must be complete when the call is made (nothing orders the forward stream after later work on the
// engine's stream -- that is the point).
// With host buffers (h_in / h_out, pinned; bfir_run_async_quad) the raw blocks travel through the ring of staging slots on
// the group's two copy streams: H2D of block b after the forward transform that last read its slot, forward b after its
// H2D (per-block events, so the first transform starts when the first block has arrived), inverse b after the D2H that
// last drained its slot, D2H b after inverse b; the ticket events ride on the output-copy stream.
int Engine::staged_blocks(int nb, const void *const *d_in, void *const *d_out, const void *const *h_in, void *const *h_out)
{
    int rc;
    const bool host = h_in != nullptr;
    const size_t cbuf = (size_t)N * rs;
    for (int k = 0; k < 2; k++) for (int j = 0; j < nb; j++) if (!sp_acc[k][j]) BFIR_CUDA(cudaMalloc(&sp_acc[k][j], cbuf * Ct));
    Group &grp = groups[0];
    if (!sp_open || (host && !async_copies)) {
        if ((rc = close_async()) != BFIR_OK) return rc;
        BFIR_CUDA(cudaEventRecord(fork_ev, stream));
        BFIR_CUDA(cudaStreamWaitEvent(sp_fwd, fork_ev, 0));
        BFIR_CUDA(cudaStreamWaitEvent(sp_inv, fork_ev, 0));
        for (int k = 0; k < 2; k++) { BFIR_CUDA(cudaEventRecord(sp_mac_done[k], stream)); BFIR_CUDA(cudaEventRecord(sp_inv_done[k], sp_inv)); }
        if (host) {
            for (int k = 0; k < kStage; k++) if ((rc = stage_alloc(k)) != BFIR_OK) return rc;
            for (int c = 0; c < host_copy_streams; c++) {
                BFIR_CUDA(cudaStreamWaitEvent(groups[c].h2d, fork_ev, 0));
                BFIR_CUDA(cudaStreamWaitEvent(groups[c].d2h, fork_ev, 0));
            }
            for (int k = 0; k < kStage; k++) { BFIR_CUDA(cudaEventRecord(grp.in_free[k], sp_fwd)); BFIR_CUDA(cudaEventRecord(grp.out_free[k], grp.d2h)); }
            async_open = async_copies = true;
            host_staged_open = true;
        }
        sp_open = true;
        sp_pairs = 0;
    }
    const int par = (int)(sp_pairs & 1ull);
    prof_nb = nb;
    int slot[4] = { 0, 0, 0, 0 };
    const void *dev_in[4] = { nullptr, nullptr, nullptr, nullptr };
    void *dev_out[4] = { nullptr, nullptr, nullptr, nullptr };
    if (host && nb > 4) { set_error("the pinned-host stage pipeline takes at most four blocks per call"); return BFIR_ERR_INVALID; }
    if (host) {   // input copies first: they are what the call's first kernels wait for
        const int sc = kStage - kStage % nb;
        for (int b = 0; b < nb; b++) {
            slot[b] = (int)((stage_next + (unsigned long long)b) % (unsigned long long)sc);
            dev_in[b] = stage_in[slot[b]]; dev_out[b] = stage_out[slot[b]];
            // consecutive blocks alternate between the copy streams of the first host_copy_streams groups: a copy that still
            // waits for its slot does not hold up the next one, and the copy engine always finds a runnable copy queued
            cudaStream_t h2d = groups[b % host_copy_streams].h2d;
            BFIR_CUDA(cudaStreamWaitEvent(h2d, grp.in_free[slot[b]], 0));
            copy_mark(ct_h2d, h2d);
            BFIR_CUDA(cudaMemcpyAsync(stage_in[slot[b]], h_in[b], in_bytes, cudaMemcpyHostToDevice, h2d));
            copy_mark(ct_h2d, h2d);
            BFIR_CUDA(cudaEventRecord(q_in_ready[b], h2d));
        }
        stage_next += (unsigned long long)nb;
        d_in = dev_in; d_out = dev_out;
    }
    tail_ready = false;
    use_abs = true;
    // forward transforms of all blocks on the forward stream
    BFIR_CUDA(cudaStreamWaitEvent(sp_fwd, sp_mac_done[par], 0));
    stage_stream = sp_fwd;
    prof_suppress = true;
    rc = BFIR_OK;
    if (!host) {
        rc = front_blocks(nb, d_in, sp_fwd);
    } else {
        for (int b = 0; b < nb && rc == BFIR_OK; b++) {
            fwd_block_offset = b;
            BFIR_CUDA(cudaStreamWaitEvent(sp_fwd, q_in_ready[b], 0));
            rc = front_group(0, d_in[b], &grp.in_free[slot[b]], true);
        }
    }
    fwd_block_offset = 0;
    prof_suppress = false;
    stage_stream = nullptr;
    if (rc != BFIR_OK) { use_abs = false; return rc; }
    BFIR_CUDA(cudaEventRecord(sp_fwd_done[par], sp_fwd));
    // partition sum on the engine's stream
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_fwd_done[par], 0));
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_inv_done[par], 0));
    prof(0);
    prof(1);   // profiling: only the partition sum's interval (1 -> 2) means anything in this mode
    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs; m.acc = sp_acc[par][0]; m.acc_next = sp_acc[par][1];
    for (int b = 0; b < nb; b++) m.acc_multi[b] = sp_acc[par][b];
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = 0; m.part_count = P;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state; m.ch_base = 0;
    m.use_abs_block = 1; m.abs_block = host_blockcounter;
    const int split = nb == 4 ? quad_split : mac_split;
    int mthreads = nb == 4 && rs == 8 ? quad_threads : 256;
    dim3 grid((N / 8 + mthreads / split - 1) / (mthreads / split), Ct);
    mac_kernel_t mk;
    size_t msmem = 0;
    if (nb == 8) {   // eight blocks: a thread owns 2 (double) or 4 (float) reals of a group, one slice
        mthreads = 128;
        const int w = rs == 4 ? mac_oct_reals_per_thread<float>() : mac_oct_reals_per_thread<double>();
        grid = dim3((N / w + mthreads - 1) / mthreads, Ct);
        mk = rs == 4 ? mac_oct_kernel<float>() : mac_oct_kernel<double>();
        if (mac_persist_sms > 0) {   // BFIR_MAC_SMS=n: persistent CTAs alone on n SMs, the rest of the GPU left to the transforms
            mthreads = 256;
            m.tiles_x = (N / w + mthreads - 1) / mthreads; m.n_ch_launch = Ct;
            grid = dim3(mac_persist_sms, 1);
            mk = rs == 4 ? mac_oct_persistent_kernel<float>() : mac_oct_persistent_kernel<double>();
            msmem = 160 * 1024;
            static bool configured[2] = { false, false };
            if (!configured[rs == 8]) { BFIR_CUDA(cudaFuncSetAttribute((const void *)mk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem)); configured[rs == 8] = true; }
        }
    } else if (nb == 4) mk = rs == 4 ? mac_quad_kernel_for_split<float>(split) : mac_quad_kernel_for_split<double>(split, mthreads);
    else mk = rs == 4 ? mac_pair_kernel_for_split<float>(split) : mac_pair_kernel_for_split<double>(split);
    mk<<<grid, mthreads, msmem, stream>>>(m);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    prof(2);
    prof(3);
    prof_next();
    BFIR_CUDA(cudaEventRecord(sp_mac_done[par], stream));
    // inverse transforms on the inverse stream
    BFIR_CUDA(cudaStreamWaitEvent(sp_inv, sp_mac_done[par], 0));
    stage_stream = sp_inv;
    prof_suppress = true;
    if (!host) {
        rc = back_blocks(nb, sp_acc[par], d_out, sp_inv);
    } else {
        for (int b = 0; b < nb && rc == BFIR_OK; b++) {
            acc_override = sp_acc[par][b];
            BFIR_CUDA(cudaStreamWaitEvent(sp_inv, grp.out_free[slot[b]], 0));
            rc = back_group(0, d_out[b]);
            BFIR_CUDA(cudaEventRecord(q_out_ready[b], sp_inv));
        }
    }
    acc_override = nullptr;
    prof_suppress = false;
    stage_stream = nullptr;
    use_abs = false;
    if (rc != BFIR_OK) return rc;
    BFIR_CUDA(cudaEventRecord(sp_inv_done[par], sp_inv));
    if (host) {
        for (int b = 0; b < nb; b++) {
            cudaStream_t d2h = groups[b % host_copy_streams].d2h;
            BFIR_CUDA(cudaStreamWaitEvent(d2h, q_out_ready[b], 0));
            copy_mark(ct_d2h, d2h);
            BFIR_CUDA(cudaMemcpyAsync(h_out[b], stage_out[slot[b]], out_bytes, cudaMemcpyDeviceToHost, d2h));
            copy_mark(ct_d2h, d2h);
            BFIR_CUDA(cudaEventRecord(grp.out_free[slot[b]], d2h));
            const int tslot = (int)((next_ticket + b) % kMaxInflight);
            if (ticket_ev[tslot][0] == nullptr) BFIR_CUDA(cudaEventCreateWithFlags(&ticket_ev[tslot][0], cudaEventDisableTiming));
            BFIR_CUDA(cudaEventRecord(ticket_ev[tslot][0], d2h));
        }
    }
    sp_pairs++;
    for (int b = 0; b < nb; b++) finish_block();
    return BFIR_OK;
}

// four consecutive blocks of pinned host buffers through the stage pipeline; returns the ticket of the fourth block.
// Needs one stream group and the steady state; otherwise two two-block calls.
long long Engine::run_host_async_quad(const void *const in[4], void *const out[4])
{
    int rc;
    if (!pair_ok() || n_groups != 1 || !staged_enabled) {
        const long long t0 = run_host_async_pair(in[0], in[1], out[0], out[1]);
        if (t0 < 0) return t0;
        return run_host_async_pair(in[2], in[3], out[2], out[3]);
    }
    if (next_ticket + 3 - done_ticket >= kMaxInflight && (rc = wait_ticket(next_ticket + 3 - kMaxInflight)) != BFIR_OK) return rc;
    if ((rc = staged_blocks(4, nullptr, nullptr, in, out)) != BFIR_OK) return rc;
    next_ticket += 4;
    return next_ticket - 1;
}

int Engine::staged_pair(const void *d_in0, const void *d_in1, void *d_out0, void *d_out1)
{
    const void *in[2] = { d_in0, d_in1 };
    void *out[2] = { d_out0, d_out1 };
    return staged_blocks(2, in, out);
}

// Partition shard with the fused reduce, FOUR blocks per call (steady state: every partition live). run_partial_quad:
// four forward transforms (+ input crossbar), ONE four-block partition sum over this rank's partitions (a shard of
// P/G partitions costs (2P/G + 7) spectra per channel for four blocks instead of 4 (2P/G + 1)), the four partial
// results pushed to their owners (straight from the partition-sum kernel, or from the output crossbar), then the
// arrival flag. run_finish_quad: wait for every source rank's flag, then per block sum the own channels' slots and
// run the output stage on them. No collective and no host synchronisation between the two.
int Engine::run_partial_quad(const void *const d_in[4])
{
    if (!peer.enabled || !peer_ready()) { set_error("run_partial_quad needs a connected peer shard"); return BFIR_ERR_INVALID; }
    if (peer_quad_pending) { set_error("run_partial_quad called twice without run_finish_quad"); return BFIR_ERR_INVALID; }
    if (host_blockcounter < (unsigned int)P || xfade_pending || n_groups != 1) { set_error("four-block shard calls need the delay line filled (%d blocks) and no pending filter swap", P); return BFIR_ERR_NOT_READY; }
    const size_t cbuf = (size_t)N * rs;
    if (!acc_pair) BFIR_CUDA(cudaMalloc(&acc_pair, cbuf * Ct));
    for (int k = 0; k < 2; k++) if (!acc_quad[k]) BFIR_CUDA(cudaMalloc(&acc_quad[k], cbuf * Ct));
    int rc = BFIR_OK;
    prof_nb = 4;
    prof(0);
    prof_suppress = true;
    for (int b = 0; b < 4 && rc == BFIR_OK; b++) { fwd_block_offset = b; rc = front_group(0, d_in[b], nullptr, true); }
    fwd_block_offset = 0;
    prof_suppress = false;
    if (rc != BFIR_OK) return rc;
    prof(1);
    tail_ready = false;
    const int base = 2 + (int)(peer_epoch & 1u) * 8;
    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs;
    m.acc_multi[0] = acc; m.acc_multi[1] = acc_pair; m.acc_multi[2] = acc_quad[0]; m.acc_multi[3] = acc_quad[1];
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = part_begin; m.part_count = part_count;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state; m.block_offset = 0; m.ch_base = 0;
    if (!xbar) { m.push = peer; m.push_phase = base; }
    const int mthreads = rs == 8 ? quad_threads : 256;
    // slices: a shard has few partitions per channel, so split less than the whole-filter heuristic would
    int split = quad_split;
    while (split > 1 && split * 4 > part_count) split >>= 1;
    dim3 grid((N / 8 + mthreads / split - 1) / (mthreads / split), Ct);
    mac_kernel_t mk = rs == 4 ? mac_quad_kernel_for_split<float>(split) : mac_quad_kernel_for_split<double>(split, mthreads);
    mk<<<grid, mthreads, 0, stream>>>(m);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    prof(2);
    if (xbar) {
        for (int b = 0; b < 4; b++) {   // partial out-mix of block b, rows pushed to their owners
            XbarArgs x = {};
            x.in = m.acc_multi[b]; x.in_stride = N; x.out = yacc; x.out_stride = N; x.slot_stride = 0;
            x.gains = gains_out; x.n_in = C; x.n_out = Co; x.N = N; x.n_streams = S; x.stream_base = 0;
            x.state = nullptr; x.n_slots = Pslots; x.n_parts = P; x.push = peer; x.push_state = state; x.push_phase = base + b;
            xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(C) : xbar_kernel_for<double>(C);
            xk<<<dim3((N + 255) / 256, S), 256, (size_t)Co * C * rs, stream>>>(x);
            count_launch();
        }
        BFIR_CUDA(cudaGetLastError());
    }
    peer_epoch++;
    peer_signal_kernel<<<1, 32, 0, stream>>>(peer, peer_epoch, peer.flag_offset);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    peer_quad_pending = true;
    return BFIR_OK;
}

int Engine::run_finish_quad(void *const d_out[4])
{
    if (!peer_quad_pending) { set_error("run_finish_quad without run_partial_quad"); return BFIR_ERR_INVALID; }
    peer_quad_pending = false;
    peer_wait_kernel<<<1, 32, 0, stream>>>(peer, peer_epoch, d_peer_timeout, peer.flag_offset);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    const int base = 2 + (int)((peer_epoch - 1u) & 1u) * 8;
    int rc = BFIR_OK;
    prof_suppress = true;
    for (int b = 0; b < 4 && rc == BFIR_OK; b++) { peer_phase = base + b; rc = back_group(0, d_out[b]); }
    peer_phase = -1;
    prof_suppress = false;
    prof(3);
    prof_next();
    for (int b = 0; b < 4; b++) finish_block();
    return rc;
}

// The four-block shard call through the stage pipeline: call k = blocks t .. t+3 of a peer-connected partition shard.
//   forward stream:  after the partition sum of call k-2 (slot reuse, as in staged_blocks): four forward transforms
//                    (+ input crossbar)
//   engine's stream: after those and after this rank has SEEN every peer's flag of call k-1: ONE four-block partition
//                    sum over the rank's partitions, the pushes (from the sum or from the output crossbar); then,
//                    after this rank's own output stage of call k-1 has finished, the arrival flag of call k
//   inverse stream:  after the flag went out: wait for every source rank's flag of call k, then per block sum the
//                    own channels' slots and run the output stage
// Receive-buffer phases are reused every second call. A peer overwrites this rank's phase set of call k-2 with its
// pushes of call k only after it has seen this rank's flag of call k-1, and that flag is raised only after this
// rank's output stage of call k-2 has read the set: no collective, no host synchronisation, and the three stages of
// neighbouring calls overlap on every rank.
int Engine::shard_blocks_staged(int nb, const void *const *d_in, void *const *d_out)
{
    if (!peer.enabled || !peer_ready()) { set_error("shard_quad_staged needs a connected peer shard"); return BFIR_ERR_INVALID; }
    if (peer_quad_pending) { set_error("run_partial_quad is pending: finish it first"); return BFIR_ERR_INVALID; }
    if (host_blockcounter < (unsigned int)P || xfade_pending || n_groups != 1) { set_error("four-block shard calls need the delay line filled (%d blocks) and no pending filter swap", P); return BFIR_ERR_NOT_READY; }
    const size_t cbuf = (size_t)N * rs;
    if (!acc_pair) BFIR_CUDA(cudaMalloc(&acc_pair, cbuf * Ct));
    for (int k = 0; k < 2; k++) if (!acc_quad[k]) BFIR_CUDA(cudaMalloc(&acc_quad[k], cbuf * Ct));
    if (nb == 8) for (int k = 0; k < 4; k++) if (!acc_oct[k]) BFIR_CUDA(cudaMalloc(&acc_oct[k], cbuf * Ct));
    int rc;
    if (!sp_open) {
        if ((rc = close_async()) != BFIR_OK) return rc;
        BFIR_CUDA(cudaEventRecord(fork_ev, stream));
        BFIR_CUDA(cudaStreamWaitEvent(sp_fwd, fork_ev, 0));
        BFIR_CUDA(cudaStreamWaitEvent(sp_inv, fork_ev, 0));
        for (int k = 0; k < 2; k++) {
            BFIR_CUDA(cudaEventRecord(sp_mac_done[k], stream));
            BFIR_CUDA(cudaEventRecord(sp_inv_done[k], sp_inv));
            BFIR_CUDA(cudaEventRecord(sp_arrived[k], sp_inv));
        }
        sp_open = true;
        sp_pairs = 0;
    }
    const int par = (int)(sp_pairs & 1ull);
    prof_nb = nb;
    tail_ready = false;
    use_abs = true;
    // forward stage
    BFIR_CUDA(cudaStreamWaitEvent(sp_fwd, sp_mac_done[par], 0));
    st_mark(sp_fwd);
    rc = BFIR_OK;
    if (xbar && peer.xin_offset != 0) {
        // SHARDED input stage (BFIR_SHARD_INPUTS=1): this rank transforms only its own inputs (own_in_count x nb CTAs
        // instead of Ci x nb), stores the spectra into every peer's input region -- phase (call parity, block) -- and
        // raises its input flag; once every rank's flag of this call is in, the whole input crossbar runs on the
        // gathered spectra. All on the forward stream, so a rank's input flag of call k+1 also says that its crossbar of
        // call k has read phase set k & 1, which its peers overwrite with call k+2 only after they have seen that flag.
        const int ibase = (int)(peer_in_epoch & 1u) * 8;
        char *xin_all = (char *)recv + peer.xin_offset;
        const long long row_bytes = (long long)N * rs;
        const unsigned int t0 = host_blockcounter;
        if (!plan8) BFIR_CUDA(cudaMalloc(&plan8, (size_t)8 * Cit * L * rs));
        PlanarArgs pa = {};
        for (int bb = 0; bb < nb; bb++) pa.raw[bb] = d_in[bb];
        pa.plan = plan8; pa.L = L; pa.n_ch = Cit; pa.fmt = in_sf.format; pa.nb = nb;
        if (rs == 4) raw_to_planar_kernel<float><<<dim3((L + 255) / 256, Cit, nb), 256, 0, sp_fwd>>>(pa);
        else raw_to_planar_kernel<double><<<dim3((L + 255) / 256, Cit, nb), 256, 0, sp_fwd>>>(pa);
        count_launch();
        if (own_in_count > 0) {
            FwdArgs f = {};
            f.in_mode = IN_PLANAR2; f.out_layout = LAYOUT_ORD;
            f.scale_in = 1.0; f.scale_out = in_sf.scale;
            f.out = xin_all + (long long)ibase * Cit * row_bytes; f.out_stride_x = N; f.out_stride_y = (long long)Cit * N;
            f.n_channels = Cit; f.ch_per_stream = Ci; f.fmt = in_sf.format; f.ch_base = own_in_first;
            for (int bb = 0; bb < nb; bb++) {
                f.hi_multi[bb] = (char *)plan8 + (size_t)bb * Cit * L * rs;
                f.lo_multi[bb] = bb == 0 ? (const void *)((char *)prev + ((size_t)(t0 & 1u) * Cit) * L * rs) : (const void *)((char *)plan8 + (size_t)(bb - 1) * Cit * L * rs);
            }
            cudaError_t e = launch_rfft_forward(rs, log2m, fft_r0, dim3(own_in_count, nb), sp_fwd, f, tw);
            count_launch();
            if (e != cudaSuccess) { set_error("forward launch failed: %s", cudaGetErrorString(e)); use_abs = false; return BFIR_ERR_CUDA; }
        }
        BFIR_CUDA(cudaMemcpyAsync((char *)prev + ((size_t)(((t0 + (unsigned int)nb - 1u) & 1u) ^ 1u) * Cit) * L * rs,
                                  (char *)plan8 + (size_t)(nb - 1) * Cit * L * rs, (size_t)Cit * L * rs, cudaMemcpyDeviceToDevice, sp_fwd));
        if (own_in_count > 0) {
            peer_bcast_kernel<<<dim3((unsigned)((row_bytes / 16 + 255) / 256), own_in_count * nb, peer.world - 1), 256, 0, sp_fwd>>>(peer, ibase, own_in_first, row_bytes, own_in_count);
            count_launch();
        }
        peer_in_epoch++;
        peer_signal_kernel<<<1, 32, 0, sp_fwd>>>(peer, peer_in_epoch, peer.flag_in_offset);
        peer_wait_kernel<<<1, 32, 0, sp_fwd>>>(peer, peer_in_epoch, d_peer_timeout, peer.flag_in_offset);
        count_launch(2);
        XbarArgs x = {};
        x.in_stride = N; x.out = fdl; x.out_stride = (long long)Pslots * N; x.slot_stride = N;
        x.gains = gains_in; x.n_in = Ci; x.n_out = C; x.N = N; x.n_streams = S; x.stream_base = 0;
        x.state = state; x.n_slots = Pslots; x.n_parts = P; x.slot_offset = 0; x.procblocks = procblocks; x.pb_inc = pb_inc;
        x.use_abs_block = 1; x.abs_block = t0;
        x.n_multi = nb;
        for (int bb = 0; bb < nb; bb++) { x.in_multi[bb] = xin_all + (long long)(ibase + bb) * Cit * row_bytes; x.out_multi[bb] = nullptr; }
        xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(Ci) : xbar_kernel_for<double>(Ci);
        xk<<<dim3((N + 255) / 256, S, nb), 256, (size_t)C * Ci * rs, sp_fwd>>>(x);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    } else {
        prof_suppress = true;
        rc = front_blocks(nb, d_in, sp_fwd);
        prof_suppress = false;
    }
    if (rc != BFIR_OK) { use_abs = false; return rc; }
    st_mark(sp_fwd);
    BFIR_CUDA(cudaEventRecord(sp_fwd_done[par], sp_fwd));
    // partition sum + pushes on the engine's stream
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_fwd_done[par], 0));
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_arrived[par ^ 1], 0));
    st_mark(stream);
    prof(0);
    prof(1);
    const int base = 2 + (int)(peer_epoch & 1u) * 8;
    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs;
    m.acc_multi[0] = acc; m.acc_multi[1] = acc_pair; m.acc_multi[2] = acc_quad[0]; m.acc_multi[3] = acc_quad[1];
    for (int k = 0; k < 4; k++) m.acc_multi[4 + k] = acc_oct[k];
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = part_begin; m.part_count = part_count;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state; m.block_offset = 0; m.ch_base = 0;
    m.use_abs_block = 1; m.abs_block = host_blockcounter;
    if (!xbar) { m.push = peer; m.push_phase = base; }
    int mthreads = rs == 8 ? quad_threads : 256;
    int split = quad_split;
    while (split > 1 && split * 4 > part_count) split >>= 1;
    dim3 grid((N / 8 + mthreads / split - 1) / (mthreads / split), Ct);
    mac_kernel_t mk = rs == 4 ? mac_quad_kernel_for_split<float>(split) : mac_quad_kernel_for_split<double>(split, mthreads);
    if (nb == 8) {
        mthreads = 128;
        const int w = rs == 4 ? mac_oct_reals_per_thread<float>() : mac_oct_reals_per_thread<double>();
        grid = dim3((N / w + mthreads - 1) / mthreads, Ct);
        mk = rs == 4 ? mac_oct_kernel<float>() : mac_oct_kernel<double>();
    }
    mk<<<grid, mthreads, 0, stream>>>(m);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    prof(2);
    st_mark(stream);
    if (xbar && batch_stages) {   // partial out-mix of all nb blocks in one launch, rows pushed to their owners
        XbarArgs x = {};
        x.in_stride = N; x.out = yacc; x.out_stride = N; x.slot_stride = 0;
        x.gains = gains_out; x.n_in = C; x.n_out = Co; x.N = N; x.n_streams = S; x.stream_base = 0;
        x.state = nullptr; x.n_slots = Pslots; x.n_parts = P; x.push = peer; x.push_state = state; x.push_phase = base;
        x.n_multi = nb;
        for (int b = 0; b < nb; b++) { x.in_multi[b] = m.acc_multi[b]; x.out_multi[b] = nullptr; }
        xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(C) : xbar_kernel_for<double>(C);
        xk<<<dim3((N + 255) / 256, S, nb), 256, (size_t)Co * C * rs, stream>>>(x);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    } else if (xbar) {
        for (int b = 0; b < nb; b++) {
            XbarArgs x = {};
            x.in = m.acc_multi[b]; x.in_stride = N; x.out = yacc; x.out_stride = N; x.slot_stride = 0;
            x.gains = gains_out; x.n_in = C; x.n_out = Co; x.N = N; x.n_streams = S; x.stream_base = 0;
            x.state = nullptr; x.n_slots = Pslots; x.n_parts = P; x.push = peer; x.push_state = state; x.push_phase = base + b;
            xbar_kernel_t xk = rs == 4 ? xbar_kernel_for<float>(C) : xbar_kernel_for<double>(C);
            xk<<<dim3((N + 255) / 256, S), 256, (size_t)Co * C * rs, stream>>>(x);
            count_launch();
        }
        BFIR_CUDA(cudaGetLastError());
    }
    prof(3);
    prof_next();
    st_mark(stream);
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_inv_done[par ^ 1], 0));   // own output stage of call k-1 has read its phase set
    peer_epoch++;
    peer_signal_kernel<<<1, 32, 0, stream>>>(peer, peer_epoch, peer.flag_offset);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    st_mark(stream);
    BFIR_CUDA(cudaEventRecord(sp_mac_done[par], stream));
    // output stage on the inverse stream
    BFIR_CUDA(cudaStreamWaitEvent(sp_inv, sp_mac_done[par], 0));
    peer_wait_kernel<<<1, 32, 0, sp_inv>>>(peer, peer_epoch, d_peer_timeout, peer.flag_offset);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    st_mark(sp_inv);
    BFIR_CUDA(cudaEventRecord(sp_arrived[par], sp_inv));
    prof_suppress = true;
    peer_phase = base;
    rc = back_blocks(nb, nullptr, d_out, sp_inv);
    peer_phase = -1;
    prof_suppress = false;
    use_abs = false;
    if (rc != BFIR_OK) return rc;
    st_mark(sp_inv);
    BFIR_CUDA(cudaEventRecord(sp_inv_done[par], sp_inv));
    sp_pairs++;
    for (int b = 0; b < nb; b++) finish_block();
    return BFIR_OK;
}

// four consecutive blocks of one group with ONE partition-sum launch (single precision): see partition_mac_multi_kernel
int Engine::quad_group(int g, const void *const d_in[4], void *const d_out[4])
{
    const Group &grp = groups[g];
    const int s0 = n_groups == 1 ? 0 : grp.s0, ns = n_groups == 1 ? S : grp.s1 - grp.s0;
    int rc = BFIR_OK;
    if (g == 0) prof_nb = 4;
    if (g == 0) prof(0);
    prof_suppress = true;
    for (int b = 0; b < 4 && rc == BFIR_OK; b++) { fwd_block_offset = b; rc = front_group(g, d_in[b], nullptr, true); }
    fwd_block_offset = 0;
    prof_suppress = false;
    if (rc != BFIR_OK) return rc;
    if (g == 0) prof(1);
    tail_ready = false;
    MacArgs m = {};
    m.fdl = fdl; m.coeffs = coeffs;
    m.acc_multi[0] = acc; m.acc_multi[1] = acc_pair; m.acc_multi[2] = acc_quad[0]; m.acc_multi[3] = acc_quad[1];
    m.fdl_stride_ch = (long long)Pslots * N; m.coeff_stride_ch = (long long)coeff_alloc * N;
    m.N = N; m.n_slots = Pslots; m.n_parts = P; m.part_begin = 0; m.part_count = P;
    m.coeff_blocks = coeff_blocks; m.coeff_map = coeff_map; m.procblocks = procblocks; m.state = state + g; m.block_offset = 0; m.ch_base = s0 * C;
    const int qsplit = quad_split, mthreads = rs == 8 ? quad_threads : 256;
    dim3 grid((N / 8 + mthreads / qsplit - 1) / (mthreads / qsplit), ns * C);
    mac_kernel_t mk = rs == 4 ? mac_quad_kernel_for_split<float>(qsplit) : mac_quad_kernel_for_split<double>(qsplit, mthreads);
    mk<<<grid, mthreads, 0, gstream(g)>>>(m);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    if (g == 0) prof(2);
    prof_suppress = true;
    for (int b = 0; b < 4 && rc == BFIR_OK; b++) { acc_override = b == 0 ? nullptr : m.acc_multi[b]; rc = back_group(g, d_out[b]); }
    acc_override = nullptr;
    prof_suppress = false;
    if (g == 0) { prof(3); prof_next(); }
    return rc;
}

// four consecutive blocks on device buffers (joined like bfir_run_device). Single precision, steady state, no crossbar;
// otherwise two pair steps.
int Engine::enqueue_quad(const void *const d_in[4], void *const d_out[4], bool staged)
{
    int rc;
    if (!pair_ok()) {
        rc = enqueue_pair(d_in[0], d_in[1], d_out[0], d_out[1], staged, staged);
        if (rc == BFIR_OK) rc = enqueue_pair(d_in[2], d_in[3], d_out[2], d_out[3], staged, staged);
        return rc;
    }
    if (staged && n_groups == 1 && staged_enabled) return staged_blocks(4, d_in, d_out);
    if ((rc = close_staged()) != BFIR_OK) return rc;
    const size_t cbuf = (size_t)N * rs;
    if (!acc_pair) BFIR_CUDA(cudaMalloc(&acc_pair, cbuf * Ct));
    for (int k = 0; k < 2; k++) if (!acc_quad[k]) BFIR_CUDA(cudaMalloc(&acc_quad[k], cbuf * Ct));
    if ((rc = close_async()) != BFIR_OK) return rc;
    if ((rc = fork()) != BFIR_OK) return rc;
    for (int g = 0; g < n_groups && rc == BFIR_OK; g++) rc = quad_group(g, d_in, d_out);
    for (int b = 0; b < 4; b++) finish_block();
    if (rc == BFIR_OK) rc = join();
    return rc;
}

// eight consecutive blocks on device buffers through the stage pipeline (joined at the end unless `staged`): ONE
// eight-block partition-sum launch. One stream group, steady state, enough work for a one-slice kernel; otherwise two
// four-block calls.
int Engine::enqueue_oct(const void *const d_in[8], void *const d_out[8], bool staged)
{
    int rc;
    const long long threads_needed = (long long)Ct * (N / (rs == 4 ? 8 : 4));
    if (!pair_ok() || n_groups != 1 || !staged_enabled || threads_needed < 148LL * 128 * 4) {
        rc = enqueue_quad(d_in, d_out, staged);
        if (rc == BFIR_OK) rc = enqueue_quad(d_in + 4, d_out + 4, staged);
        return rc;
    }
    rc = staged_blocks(8, d_in, d_out);
    if (rc == BFIR_OK && !staged) rc = close_async();
    return rc;
}

// two consecutive blocks on device buffers. Falls back to two single-block steps while the delay line is still
// filling, on a partition shard and with a pending filter swap.
int Engine::enqueue_pair(const void *d_in0, const void *d_in1, void *d_out0, void *d_out1, bool pipelined, bool staged)
{
    int rc;
    if (!pair_ok()) {
        rc = pipelined ? enqueue_block_pipelined(d_in0, d_out0) : enqueue_block(d_in0, d_out0);
        if (rc == BFIR_OK) rc = pipelined ? enqueue_block_pipelined(d_in1, d_out1) : enqueue_block(d_in1, d_out1);
        return rc;
    }
    if (staged && n_groups == 1 && staged_enabled) return staged_pair(d_in0, d_in1, d_out0, d_out1);
    if (!acc_pair) BFIR_CUDA(cudaMalloc(&acc_pair, (size_t)N * rs * Ct));
    if ((rc = close_staged()) != BFIR_OK) return rc;
    if (!pipelined) { if ((rc = close_async()) != BFIR_OK) return rc; }
    if (!async_open) {
        if ((rc = join_tail()) != BFIR_OK || (rc = fork()) != BFIR_OK) return rc;
        if (pipelined) async_open = n_groups > 1;
    }
    rc = BFIR_OK;
    for (int g = 0; g < n_groups && rc == BFIR_OK; g++) rc = pair_group(g, d_in0, d_in1, d_out0, d_out1, nullptr, nullptr);
    finish_block();
    finish_block();
    if (rc == BFIR_OK && !pipelined) rc = join();
    return rc;
}

int Engine::enqueue_front(const void *d_inbuf)
{
    int rc = fork();
    for (int g = 0; g < n_groups && rc == BFIR_OK; g++) rc = front_group(g, d_inbuf);
    if (rc == BFIR_OK) rc = join();
    return rc;
}

int Engine::enqueue_back(void *d_outbuf)
{
    int rc = fork();
    for (int g = 0; g < n_groups && rc == BFIR_OK; g++) rc = back_group(g, d_outbuf);
    if (rc == BFIR_OK) rc = join();
    finish_block();
    return rc;
}

// one whole block step on device buffers: every group runs front and back on its own stream
int Engine::enqueue_block(const void *d_inbuf, void *d_outbuf)
{
    int rc = close_async();
    if (rc == BFIR_OK) rc = fork();
    for (int g = 0; g < n_groups && rc == BFIR_OK; g++) {
        rc = front_group(g, d_inbuf);
        if (rc == BFIR_OK) rc = back_group(g, d_outbuf);
    }
    if (rc == BFIR_OK) rc = join();
    finish_block();
    return rc;
}

// the same without the join: the groups of consecutive blocks run into each other (a group that is done with
// block t starts block t+1 while others still work on t); close_async() joins
int Engine::enqueue_block_pipelined(const void *d_inbuf, void *d_outbuf)
{
    int rc = close_staged();
    if (rc != BFIR_OK) return rc;
    if (!async_open) {
        if ((rc = join_tail()) != BFIR_OK || (rc = fork()) != BFIR_OK) return rc;
        async_open = n_groups > 1;
    }
    for (int g = 0; g < n_groups && rc == BFIR_OK; g++) {
        rc = front_group(g, d_inbuf);
        if (rc == BFIR_OK) rc = back_group(g, d_outbuf);
    }
    finish_block();
    return rc;
}

// wait for the stream and apply the reference's NaN/Inf abort (brutefir.cpp:316-321)
int Engine::sync_and_probe(bool allow_rollback, cudaEvent_t wait_for)
{
    if (wait_for) {   // bfir_run: the output is home, look-ahead work may still run (and stays un-joined)
        BFIR_CUDA(cudaEventSynchronize(wait_for));
    } else {
        int arc = close_async();
        if (arc != BFIR_OK) return arc;
        BFIR_CUDA(cudaStreamSynchronize(stream));
    }
    done_ticket = next_ticket;
    prof_collect();
    if (peer.enabled && d_peer_timeout != nullptr && peer_epoch > 0) {   // four-block shard calls: did every source rank arrive?
        int timed_out = 0;
        BFIR_CUDA(cudaMemcpy(&timed_out, d_peer_timeout, sizeof(int), cudaMemcpyDeviceToHost));
        if (timed_out) {
            cudaMemset(d_peer_timeout, 0, sizeof(int));
            set_error("partition shard: a peer rank never signalled its partial sums (time-out in the arrival wait)");
            return BFIR_ERR_CUDA;
        }
    }
    const unsigned long long n = blocks_since_sync;
    blocks_since_sync = 0;
    if (*(volatile int *)h_flag == 0) return BFIR_OK;          // nothing raised the flag: no copy needed
    *h_flag = 0;
    { int trc = join_tail(); if (trc != BFIR_OK) return trc; }
    BFIR_CUDA(cudaMemcpyAsync(h_state, state, sizeof(EngineState) * n_groups, cudaMemcpyDeviceToHost, stream));
    BFIR_CUDA(cudaStreamSynchronize(stream));
    int bad = 0x7fffffff;
    for (int g = 0; g < n_groups; g++) if (h_state[g].first_bad_channel < bad) bad = h_state[g].first_bad_channel;
    if (bad != 0x7fffffff) {
        tail_ready = false;
        pinfo("NaN or Inf values in the system! Invalid input? Aborting.\n");
        const int threads = 256, blocks = (Ct + threads - 1) / threads;
        if (n == 1 && allow_rollback) { // exact reference semantics: the aborted block does not advance the counters
            host_blockcounter--;
            // without a crossbar channel n's input and output stages belong together and the reference never ran
            // the channels after `bad`; with one, `bad` is an OUTPUT index and every filter channel has already
            // counted the block on its input side: all of them are rolled back, as if the block had not happened
            engine_abort_fixup_kernel<<<blocks, threads, 0, stream>>>(state, n_groups, procblocks, pb_inc, Ct, xbar ? -1 : bad);
            count_launch();
        } else {
            for (int g = 0; g < n_groups; g++) h_state[g].first_bad_channel = 0x7fffffff;
            cudaMemcpyAsync(state, h_state, sizeof(EngineState) * n_groups, cudaMemcpyHostToDevice, stream);
        }
        BFIR_CUDA(cudaStreamSynchronize(stream));
        set_error("NaN or Inf values in the system (channel %d)", bad);
        return BFIR_ERR_NONFINITE;
    }
    return BFIR_OK;
}

// brutefir::run on host buffers: per group H2D -> kernels -> D2H on the group's stream, so the copies
// of one group overlap the kernels of another; returns when outbuf is complete
// capture (first use per block parity) and replay the kernels of one block step: saves the per-kernel
// launch cost on the latency path. Only pointers fixed for the engine's lifetime are baked in (the staging
// buffers, the state words); whatever can change them invalidates the graphs.
int Engine::run_step_graph(bool use_tail)
{
    if (graph_flavor != (use_tail ? 1 : 0)) { invalidate_graphs(); graph_flavor = use_tail ? 1 : 0; }
    const int par = (int)(host_blockcounter & 1u);
    if (step_graph[par] == nullptr) {
        const unsigned long long before = t_launches;
        cudaGraph_t graph = nullptr;
        BFIR_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed));
        int rc = front_group(0, d_in, nullptr, use_tail);
        if (rc == BFIR_OK) rc = back_group(0, d_out, use_tail);
        cudaError_t ce = cudaStreamEndCapture(stream, &graph);
        if (rc != BFIR_OK || ce != cudaSuccess || graph == nullptr) {
            if (graph) cudaGraphDestroy(graph);
            set_error("graph capture of the block step failed");
            return rc != BFIR_OK ? rc : BFIR_ERR_CUDA;
        }
        ce = cudaGraphInstantiate(&step_graph[par], graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { step_graph[par] = nullptr; set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); return BFIR_ERR_CUDA; }
        step_graph_launches[par] = t_launches - before;
        g_launches.fetch_sub(step_graph_launches[par]);   // nothing ran during capture
    }
    BFIR_CUDA(cudaGraphLaunch(step_graph[par], stream));
    count_launch(step_graph_launches[par]);
    return BFIR_OK;
}

// the engine's stream catches up with whatever the pipelined calls left on the group streams
// the engine's stream is ordered after the look-ahead launches that are still on their own stream
int Engine::join_tail()
{
    if (!tail_inflight) return BFIR_OK;
    tail_inflight = false;
    BFIR_CUDA(cudaEventRecord(tail_join_ev, tail_stream));
    BFIR_CUDA(cudaStreamWaitEvent(stream, tail_join_ev, 0));
    return BFIR_OK;
}

// the engine's stream is ordered after the side streams of the stage pipeline
int Engine::close_staged()
{
    if (!sp_open) return BFIR_OK;
    sp_open = false;
    BFIR_CUDA(cudaEventRecord(sp_fwd_done[0], sp_fwd));
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_fwd_done[0], 0));
    BFIR_CUDA(cudaEventRecord(sp_inv_done[0], sp_inv));
    BFIR_CUDA(cudaStreamWaitEvent(stream, sp_inv_done[0], 0));
    return BFIR_OK;
}

int Engine::close_async()
{
    int trc = join_tail();
    if (trc != BFIR_OK) return trc;
    if ((trc = close_staged()) != BFIR_OK) return trc;
    if (!async_open) return BFIR_OK;
    async_open = false;
    if (async_copies) { // the output copies are the last link of every group's chain
        async_copies = false;
        const int ncopy = host_staged_open && host_copy_streams > n_groups ? host_copy_streams : n_groups;
        host_staged_open = false;
        for (int g = 0; g < ncopy; g++) {
            BFIR_CUDA(cudaEventRecord(groups[g].copies_done, groups[g].d2h));
            BFIR_CUDA(cudaStreamWaitEvent(stream, groups[g].copies_done, 0));
            BFIR_CUDA(cudaEventRecord(groups[g].copies_done, groups[g].h2d));
            BFIR_CUDA(cudaStreamWaitEvent(stream, groups[g].copies_done, 0));
        }
    }
    if (n_groups == 1) return BFIR_OK;
    return join();
}

int Engine::stage_alloc(int k)
{
    if (k == 0) { stage_in[0] = d_in; stage_out[0] = d_out; return BFIR_OK; }
    if (!stage_in[k]) BFIR_CUDA(cudaMalloc(&stage_in[k], in_bytes));
    if (!stage_out[k]) BFIR_CUDA(cudaMalloc(&stage_out[k], out_bytes));
    return BFIR_OK;
}

// all streams of the pipelined host path start after everything queued on the engine's stream so far
int Engine::open_async_copies()
{
    if (async_open && async_copies && !sp_open) return BFIR_OK;
    int rc = close_async();
    if (rc != BFIR_OK) return rc;
    for (int k = 0; k < kStage; k++) if ((rc = stage_alloc(k)) != BFIR_OK) return rc;
    BFIR_CUDA(cudaEventRecord(fork_ev, stream));
    for (int g = 0; g < n_groups; g++) {
        if (n_groups > 1) BFIR_CUDA(cudaStreamWaitEvent(groups[g].stream, fork_ev, 0));
        BFIR_CUDA(cudaStreamWaitEvent(groups[g].h2d, fork_ev, 0));
        BFIR_CUDA(cudaStreamWaitEvent(groups[g].d2h, fork_ev, 0));
        for (int k = 0; k < kStage; k++) {
            BFIR_CUDA(cudaEventRecord(groups[g].in_free[k], gstream(g)));
            BFIR_CUDA(cudaEventRecord(groups[g].out_free[k], groups[g].d2h));
        }
    }
    async_open = async_copies = true;
    return BFIR_OK;
}

// queue H2D -> block step -> D2H of one block and return its ticket (or an error code < 0). Per group three
// streams: input copies, kernels, output copies, chained by events over a ring of kStage staging slots --
//   H2D(t) after the forward transform that last read slot t % kStage;  forward(t) after H2D(t);
//   inverse(t) after the D2H that last drained slot t % kStage;         D2H(t) after inverse(t)
// so a group's copies overlap its own kernels as well as everybody else's, and run ahead of / behind them.
long long Engine::run_host_async(const void *inbuf, void *outbuf)
{
    int rc;
    if ((rc = open_async_copies()) != BFIR_OK) return rc;
    if (next_ticket - done_ticket >= kMaxInflight && (rc = wait_ticket(next_ticket - kMaxInflight)) != BFIR_OK) return rc;
    const int slot = (int)(next_ticket % kMaxInflight);
    const int k = (int)(stage_next++ % (unsigned long long)stage_count);
    for (int g = 0; g < n_groups; g++) {
        Group &grp = groups[g];
        const size_t s0 = n_groups == 1 ? 0 : (size_t)grp.s0, ns = n_groups == 1 ? (size_t)S : (size_t)(grp.s1 - grp.s0);
        const size_t ioff = s0 * L * Ci * in_sf.bytes, ibytes = ns * L * Ci * in_sf.bytes;
        const size_t ooff = s0 * L * Co * out_sf.bytes, obytes = ns * L * Co * out_sf.bytes;
        cudaStream_t st = gstream(g);
        BFIR_CUDA(cudaStreamWaitEvent(grp.h2d, grp.in_free[k], 0));
        BFIR_CUDA(cudaMemcpyAsync((char *)stage_in[k] + ioff, (const char *)inbuf + ioff, ibytes, cudaMemcpyHostToDevice, grp.h2d));
        BFIR_CUDA(cudaEventRecord(grp.in_ready, grp.h2d));
        BFIR_CUDA(cudaStreamWaitEvent(st, grp.in_ready, 0));
        if ((rc = front_group(g, stage_in[k], &grp.in_free[k])) != BFIR_OK) return rc;
        BFIR_CUDA(cudaStreamWaitEvent(st, grp.out_free[k], 0));
        if ((rc = back_group(g, stage_out[k])) != BFIR_OK) return rc;
        BFIR_CUDA(cudaEventRecord(grp.out_ready, st));
        BFIR_CUDA(cudaStreamWaitEvent(grp.d2h, grp.out_ready, 0));
        BFIR_CUDA(cudaMemcpyAsync((char *)outbuf + ooff, (const char *)stage_out[k] + ooff, obytes, cudaMemcpyDeviceToHost, grp.d2h));
        BFIR_CUDA(cudaEventRecord(grp.out_free[k], grp.d2h));
        if (ticket_ev[slot][g] == nullptr) BFIR_CUDA(cudaEventCreateWithFlags(&ticket_ev[slot][g], cudaEventDisableTiming));
        BFIR_CUDA(cudaEventRecord(ticket_ev[slot][g], grp.d2h));
    }
    finish_block();
    return next_ticket++;
}

// two consecutive blocks of host buffers per call (see pair_group); returns the ticket of the second block
long long Engine::run_host_async_pair(const void *in0, const void *in1, void *out0, void *out1)
{
    int rc;
    if (!pair_ok()) {
        const long long t0 = run_host_async(in0, out0);
        if (t0 < 0) return t0;
        return run_host_async(in1, out1);
    }
    if (!acc_pair) BFIR_CUDA(cudaMalloc(&acc_pair, (size_t)N * rs * Ct));
    if ((rc = open_async_copies()) != BFIR_OK) return rc;
    if (next_ticket + 1 - done_ticket >= kMaxInflight && (rc = wait_ticket(next_ticket + 1 - kMaxInflight)) != BFIR_OK) return rc;
    const int slot0 = (int)(next_ticket % kMaxInflight), slot1 = (int)((next_ticket + 1) % kMaxInflight);
    const int sc = stage_count >= 2 ? (stage_count & ~1) : 2;   // pairs need an even number of slots, at least two
    const int k0 = (int)(stage_next % (unsigned long long)sc), k1 = (int)((stage_next + 1) % (unsigned long long)sc);
    stage_next += 2;
    if (whole_copies && n_groups > 1) { // one copy stream each way moves whole blocks; the groups only run kernels
        cudaStream_t h2d = groups[0].h2d, d2h = groups[0].d2h;
        for (int g = 0; g < n_groups; g++) {
            BFIR_CUDA(cudaStreamWaitEvent(h2d, groups[g].in_free[k0], 0));
            BFIR_CUDA(cudaStreamWaitEvent(h2d, groups[g].in_free[k1], 0));
        }
        BFIR_CUDA(cudaMemcpyAsync(stage_in[k0], in0, in_bytes, cudaMemcpyHostToDevice, h2d));
        BFIR_CUDA(cudaMemcpyAsync(stage_in[k1], in1, in_bytes, cudaMemcpyHostToDevice, h2d));
        BFIR_CUDA(cudaEventRecord(groups[0].in_ready, h2d));
        for (int g = 0; g < n_groups; g++) {
            Group &grp = groups[g];
            cudaStream_t st = gstream(g);
            BFIR_CUDA(cudaStreamWaitEvent(st, groups[0].in_ready, 0));
            BFIR_CUDA(cudaStreamWaitEvent(st, groups[0].out_free[k0], 0));
            if ((rc = pair_group(g, stage_in[k0], stage_in[k1], stage_out[k0], stage_out[k1], &grp.in_free[k1], &groups[0].out_free[k1])) != BFIR_OK) return rc;
            BFIR_CUDA(cudaEventRecord(grp.in_free[k0], st));
            BFIR_CUDA(cudaEventRecord(grp.out_ready, st));
            BFIR_CUDA(cudaStreamWaitEvent(d2h, grp.out_ready, 0));
        }
        BFIR_CUDA(cudaMemcpyAsync(out0, stage_out[k0], out_bytes, cudaMemcpyDeviceToHost, d2h));
        BFIR_CUDA(cudaEventRecord(groups[0].out_free[k0], d2h));
        BFIR_CUDA(cudaMemcpyAsync(out1, stage_out[k1], out_bytes, cudaMemcpyDeviceToHost, d2h));
        BFIR_CUDA(cudaEventRecord(groups[0].out_free[k1], d2h));
        for (int slot : { slot0, slot1 }) {
            if (ticket_ev[slot][0] == nullptr) BFIR_CUDA(cudaEventCreateWithFlags(&ticket_ev[slot][0], cudaEventDisableTiming));
            BFIR_CUDA(cudaEventRecord(ticket_ev[slot][0], d2h));
            for (int g = 1; g < n_groups; g++) if (ticket_ev[slot][g]) { cudaEventDestroy(ticket_ev[slot][g]); ticket_ev[slot][g] = nullptr; }
        }
        finish_block();
        finish_block();
        next_ticket += 2;
        return next_ticket - 1;
    }
    for (int g = 0; g < n_groups; g++) {
        Group &grp = groups[g];
        const size_t s0 = n_groups == 1 ? 0 : (size_t)grp.s0, ns = n_groups == 1 ? (size_t)S : (size_t)(grp.s1 - grp.s0);
        const size_t ioff = s0 * L * Ci * in_sf.bytes, ibytes = ns * L * Ci * in_sf.bytes;
        const size_t ooff = s0 * L * Co * out_sf.bytes, obytes = ns * L * Co * out_sf.bytes;
        cudaStream_t st = gstream(g);
        BFIR_CUDA(cudaStreamWaitEvent(grp.h2d, grp.in_free[k0], 0));
        BFIR_CUDA(cudaStreamWaitEvent(grp.h2d, grp.in_free[k1], 0));
        BFIR_CUDA(cudaMemcpyAsync((char *)stage_in[k0] + ioff, (const char *)in0 + ioff, ibytes, cudaMemcpyHostToDevice, grp.h2d));
        BFIR_CUDA(cudaMemcpyAsync((char *)stage_in[k1] + ioff, (const char *)in1 + ioff, ibytes, cudaMemcpyHostToDevice, grp.h2d));
        BFIR_CUDA(cudaEventRecord(grp.in_ready, grp.h2d));
        BFIR_CUDA(cudaStreamWaitEvent(st, grp.in_ready, 0));
        BFIR_CUDA(cudaStreamWaitEvent(st, grp.out_free[k0], 0));   // (long done: these slots were drained kStage blocks ago)
        if ((rc = pair_group(g, stage_in[k0], stage_in[k1], stage_out[k0], stage_out[k1], &grp.in_free[k1], &grp.out_free[k1])) != BFIR_OK) return rc;
        BFIR_CUDA(cudaEventRecord(grp.in_free[k0], st));
        BFIR_CUDA(cudaEventRecord(grp.out_ready, st));
        BFIR_CUDA(cudaStreamWaitEvent(grp.d2h, grp.out_ready, 0));
        BFIR_CUDA(cudaMemcpyAsync((char *)out0 + ooff, (const char *)stage_out[k0] + ooff, obytes, cudaMemcpyDeviceToHost, grp.d2h));
        BFIR_CUDA(cudaEventRecord(grp.out_free[k0], grp.d2h));
        BFIR_CUDA(cudaMemcpyAsync((char *)out1 + ooff, (const char *)stage_out[k1] + ooff, obytes, cudaMemcpyDeviceToHost, grp.d2h));
        BFIR_CUDA(cudaEventRecord(grp.out_free[k1], grp.d2h));
        for (int slot : { slot0, slot1 }) {
            if (ticket_ev[slot][g] == nullptr) BFIR_CUDA(cudaEventCreateWithFlags(&ticket_ev[slot][g], cudaEventDisableTiming));
            BFIR_CUDA(cudaEventRecord(ticket_ev[slot][g], grp.d2h));
        }
    }
    finish_block();
    finish_block();
    next_ticket += 2;
    return next_ticket - 1;
}

// returns when the step with ticket t (and every earlier one) has delivered its output block
int Engine::wait_ticket(long long t)
{
    if (t < 0 || t >= next_ticket) { set_error("bfir_wait: unknown ticket %lld", t); return BFIR_ERR_INVALID; }
    for (; done_ticket <= t; done_ticket++) {
        const int slot = (int)(done_ticket % kMaxInflight);
        for (int g = 0; g < n_groups; g++) if (ticket_ev[slot][g]) BFIR_CUDA(cudaEventSynchronize(ticket_ev[slot][g]));
    }
    if (*(volatile int *)h_flag != 0) return sync_and_probe(false);   // a probe fired: drain everything and report
    return BFIR_OK;
}

int Engine::run_host(const void *inbuf, void *outbuf)
{
    if (async_open || sp_open) { const int arc = close_async(); if (arc != BFIR_OK) return arc; }
    const bool use_tail = tail_ready && lookahead_ok();   // acc already holds partitions 1 .. P-1 of this block
    const bool swap_block = xfade_pending;                // filters are being swapped: more swaps may follow, no look-ahead
    tail_ready = false;
    int rc;
    if (!use_tail && (rc = join_tail()) != BFIR_OK) return rc;   // a dropped look-ahead result must not race the full sum
    // latency path: one group, kernels replayed from a graph (after two plain blocks have configured them)
    if (graphs_enabled && n_groups == 1 && !xfade_pending && pcap == 0 && !peer.enabled && host_blockcounter >= 2) {
        BFIR_CUDA(cudaMemcpyAsync(d_in, inbuf, in_bytes, cudaMemcpyHostToDevice, stream));
        if ((rc = run_step_graph(use_tail)) != BFIR_OK) return rc;
        BFIR_CUDA(cudaMemcpyAsync(outbuf, d_out, out_bytes, cudaMemcpyDeviceToHost, stream));
        finish_block();
    } else {
        if ((rc = fork()) != BFIR_OK) return rc;
        for (int g = 0; g < n_groups; g++) {
            const Group &grp = groups[g];
            const size_t s0 = n_groups == 1 ? 0 : (size_t)grp.s0, ns = n_groups == 1 ? (size_t)S : (size_t)(grp.s1 - grp.s0);
            const size_t ioff = s0 * L * Ci * in_sf.bytes, ibytes = ns * L * Ci * in_sf.bytes;
            BFIR_CUDA(cudaMemcpyAsync((char *)d_in + ioff, (const char *)inbuf + ioff, ibytes, cudaMemcpyHostToDevice, gstream(g)));
            if (use_tail && n_groups > 1) BFIR_CUDA(cudaStreamWaitEvent(gstream(g), tail_done[g], 0));   // this group's look-ahead sum is in
            if ((rc = front_group(g, d_in, nullptr, use_tail)) != BFIR_OK) return rc;
        }
        for (int g = 0; g < n_groups; g++) {
            const Group &grp = groups[g];
            const size_t s0 = n_groups == 1 ? 0 : (size_t)grp.s0, ns = n_groups == 1 ? (size_t)S : (size_t)(grp.s1 - grp.s0);
            const size_t ooff = s0 * L * Co * out_sf.bytes, obytes = ns * L * Co * out_sf.bytes;
            if ((rc = back_group(g, d_out, use_tail)) != BFIR_OK) return rc;
            BFIR_CUDA(cudaMemcpyAsync((char *)outbuf + ooff, (const char *)d_out + ooff, obytes, cudaMemcpyDeviceToHost, gstream(g)));
        }
        finish_block();
        if ((rc = join()) != BFIR_OK) return rc;
    }
    BFIR_CUDA(cudaEventRecord(out_done, stream));
    if (lookahead_ok() && !swap_block) { // behind the output copies: the next block's partitions 1 .. P-1, while the host is away
        if (n_groups == 1) {
            if ((rc = tail_group(0, stream)) != BFIR_OK) return rc;
        } else {
            BFIR_CUDA(cudaStreamWaitEvent(tail_stream, out_done, 0));
            for (int g = 0; g < n_groups; g++) {
                if ((rc = tail_group(g, tail_stream)) != BFIR_OK) return rc;
                BFIR_CUDA(cudaEventRecord(tail_done[g], tail_stream));
            }
            tail_inflight = true;
        }
        tail_ready = true;
    }
    rc = sync_and_probe(true, out_done);
    return rc;
}

int Engine::get_overflow(int ch, bfir_overflow_t *out)
{
    if (ch < 0 || ch >= Cot || out == nullptr) return BFIR_ERR_INVALID;
    OverflowStats s;
    BFIR_CUDA(cudaStreamSynchronize(stream));
    BFIR_CUDA(cudaMemcpy(&s, stats + ch, sizeof(s), cudaMemcpyDeviceToHost));
    out->n_overflows = s.n_overflows;
    out->intlargest = s.intlargest;
    memcpy(&out->largest, &s.largest_bits, sizeof(double));
    out->max = ovf_max;
    return BFIR_OK;
}

} // namespace bfir

using namespace bfir;

struct bfir_engine { Engine impl; };

extern "C" {

int bfir_create_ex(bfir_engine **out, const bfir_config_t *cfg)
{
    if (out == nullptr || cfg == nullptr) return BFIR_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        set_error("no CUDA device: libbfir_b200 has no CPU fallback");
        return BFIR_ERR_CUDA;
    }
    bfir_engine *e = new bfir_engine;
    const int rc = e->impl.init(*cfg);
    if (rc != BFIR_OK) { delete e; return rc; }
    *out = e;
    return BFIR_OK;
}

int bfir_create(bfir_engine **out, int filter_length, int filter_blocks, int realsize, int channels,
                int in_format, int out_format, int sampling_rate, int apply_dither)
{
    bfir_config_t c;
    memset(&c, 0, sizeof(c));
    c.filter_length = filter_length; c.filter_blocks = filter_blocks; c.realsize = realsize; c.channels = channels;
    c.in_format = in_format; c.out_format = out_format; c.sampling_rate = sampling_rate; c.apply_dither = apply_dither;
    c.n_streams = 1; c.device = -1; c.part_begin = 0; c.part_count = 0; c.n_groups = 0; c.xbar_inputs = 0; c.xbar_outputs = 0;
    return bfir_create_ex(out, &c);
}

void bfir_destroy(bfir_engine *e) { delete e; }
int bfir_is_initialized(const bfir_engine *e) { return (e != nullptr && e->impl.initialized) ? 1 : 0; }

int bfir_set_coeff(bfir_engine *e, const void *const *coeffs, int n_coeffs, int length, int coeff_blocks, double scale)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.set_coeff(coeffs, n_coeffs, length, coeff_blocks, scale);
}

int bfir_set_coeff_map(bfir_engine *e, const int *map, int n)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    Engine &g = e->impl;
    if (map != nullptr && n != g.Ct) { bfir::set_error("bfir_set_coeff_map: %d entries for %d filter channels", n, g.Ct); return BFIR_ERR_INVALID; }
    BFIR_CUDA(cudaStreamSynchronize(g.stream));
    if (map == nullptr) {
        if (g.coeff_map) { cudaFree(g.coeff_map); g.coeff_map = nullptr; }
    } else {
        for (int c = 0; c < n; c++) if (map[c] < 0 || map[c] >= g.Ct) { bfir::set_error("bfir_set_coeff_map: entry %d = %d out of range", c, map[c]); return BFIR_ERR_INVALID; }
        if (!g.coeff_map) BFIR_CUDA(cudaMalloc((void **)&g.coeff_map, sizeof(int) * g.Ct));
        BFIR_CUDA(cudaMemcpy(g.coeff_map, map, sizeof(int) * g.Ct, cudaMemcpyHostToDevice));
    }
    g.tail_ready = false;           // a look-ahead partition sum computed with the old map is not reused
    g.invalidate_graphs();
    return BFIR_OK;
}

int bfir_set_coeff_crossfade(bfir_engine *e, const void *const *coeffs, int n_coeffs, int length, int coeff_blocks, double scale)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.load_coeff(coeffs, nullptr, 0, n_coeffs, length, coeff_blocks, scale, true);
}

int bfir_set_coeff_device(bfir_engine *e, const void *d_coeffs, long long channel_stride, int n_coeffs, int length,
                          int coeff_blocks, double scale, int crossfade)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr || d_coeffs == nullptr) return BFIR_ERR_INVALID;
    return e->impl.load_coeff(nullptr, d_coeffs, channel_stride, n_coeffs, length, coeff_blocks, scale, crossfade != 0);
}

static int check_ready(bfir_engine *e)
{
    if (e == nullptr) return BFIR_ERR_INVALID;
    int cur = -1; // a host thread that has not touched CUDA yet sits on device 0: follow the engine
    if (cudaGetDevice(&cur) == cudaSuccess && cur != e->impl.device) BFIR_CUDA(cudaSetDevice(e->impl.device));
    if (!e->impl.initialized) { set_error("run before set_coeff"); return BFIR_ERR_NOT_READY; }
    if (e->impl.xbar && !e->impl.xbar_set) { set_error("run before set_crossbar"); return BFIR_ERR_NOT_READY; }
    if (!e->impl.peer_ready()) { set_error("fused reduce: not every peer receive buffer is connected"); return BFIR_ERR_NOT_READY; }
    return BFIR_OK;
}

int bfir_run(bfir_engine *e, const void *inbuf, void *outbuf)
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (inbuf == nullptr || outbuf == nullptr) return BFIR_ERR_INVALID;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run is not available on a peer-connected partition shard (run_partial / barrier / run_finish)"); return BFIR_ERR_INVALID; }
    return e->impl.run_host(inbuf, outbuf);
}

int bfir_run_device(bfir_engine *e, const void *d_inbuf, void *d_outbuf)
{
    if (e != nullptr) e->impl.close_async();
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run_device is not available on a peer-connected partition shard (run_partial / barrier / run_finish)"); return BFIR_ERR_INVALID; }
    return e->impl.enqueue_block(d_inbuf, d_outbuf);
}

int bfir_run_device_pipelined(bfir_engine *e, const void *d_inbuf, void *d_outbuf)
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run_device_pipelined is not available on a partition shard"); return BFIR_ERR_INVALID; }
    return e->impl.enqueue_block_pipelined(d_inbuf, d_outbuf);
}

int bfir_run_device_pair(bfir_engine *e, const void *d_in0, const void *d_in1, void *d_out0, void *d_out1, int pipelined)
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in0 == nullptr || d_in1 == nullptr || d_out0 == nullptr || d_out1 == nullptr) return BFIR_ERR_INVALID;
    return e->impl.enqueue_pair(d_in0, d_in1, d_out0, d_out1, pipelined != 0, pipelined == BFIR_PAIR_STAGED);
}

int bfir_run_device_quad(bfir_engine *e, const void *const d_in[4], void *const d_out[4])
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in == nullptr || d_out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 4; b++) if (d_in[b] == nullptr || d_out[b] == nullptr) return BFIR_ERR_INVALID;
    return e->impl.enqueue_quad(d_in, d_out);
}

int bfir_run_device_quad_staged(bfir_engine *e, const void *const d_in[4], void *const d_out[4])
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in == nullptr || d_out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 4; b++) if (d_in[b] == nullptr || d_out[b] == nullptr) return BFIR_ERR_INVALID;
    return e->impl.enqueue_quad(d_in, d_out, true);
}

long long bfir_run_async_pair(bfir_engine *e, const void *in0, const void *in1, void *out0, void *out1)
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (in0 == nullptr || in1 == nullptr || out0 == nullptr || out1 == nullptr) return BFIR_ERR_INVALID;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run_async_pair is not available on a partition shard"); return BFIR_ERR_INVALID; }
    return e->impl.run_host_async_pair(in0, in1, out0, out1);
}

int bfir_run_device_oct(bfir_engine *e, const void *const d_in[8], void *const d_out[8], int staged)
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in == nullptr || d_out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 8; b++) if (d_in[b] == nullptr || d_out[b] == nullptr) return BFIR_ERR_INVALID;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run_device_oct is not available on a partition shard"); return BFIR_ERR_INVALID; }
    return e->impl.enqueue_oct(d_in, d_out, staged != 0);
}

long long bfir_run_async_quad(bfir_engine *e, const void *const in[4], void *const out[4])
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (in == nullptr || out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 4; b++) if (in[b] == nullptr || out[b] == nullptr) return BFIR_ERR_INVALID;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run_async_quad is not available on a partition shard"); return BFIR_ERR_INVALID; }
    return e->impl.run_host_async_quad(in, out);
}

int bfir_join(bfir_engine *e)
{
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.close_async();
}

long long bfir_run_async(bfir_engine *e, const void *inbuf, void *outbuf)
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (inbuf == nullptr || outbuf == nullptr) return BFIR_ERR_INVALID;
    if (e->impl.peer.enabled) { bfir::set_error("bfir_run_async is not available on a partition shard"); return BFIR_ERR_INVALID; }
    return e->impl.run_host_async(inbuf, outbuf);
}

int bfir_wait(bfir_engine *e, long long ticket)
{
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.wait_ticket(ticket);
}

int bfir_run_partial_device(bfir_engine *e, const void *d_inbuf)
{
    if (e != nullptr) e->impl.close_async();
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    return e->impl.enqueue_front(d_inbuf);
}

int bfir_run_finish_device(bfir_engine *e, void *d_outbuf)
{
    if (e != nullptr) e->impl.close_async();
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    return e->impl.enqueue_back(d_outbuf);
}

int bfir_run_partial_quad_device(bfir_engine *e, const void *const d_in[4])
{
    if (e != nullptr) e->impl.close_async();
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 4; b++) if (d_in[b] == nullptr) return BFIR_ERR_INVALID;
    return e->impl.run_partial_quad(d_in);
}

int bfir_run_finish_quad_device(bfir_engine *e, void *const d_out[4])
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 4; b++) if (d_out[b] == nullptr) return BFIR_ERR_INVALID;
    return e->impl.run_finish_quad(d_out);
}

int bfir_run_shard_quad_staged(bfir_engine *e, const void *const d_in[4], void *const d_out[4])
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in == nullptr || d_out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 4; b++) if (d_in[b] == nullptr || d_out[b] == nullptr) return BFIR_ERR_INVALID;
    return e->impl.shard_blocks_staged(4, d_in, d_out);
}

int bfir_run_shard_oct_staged(bfir_engine *e, const void *const d_in[8], void *const d_out[8])
{
    int rc = check_ready(e);
    if (rc != BFIR_OK) return rc;
    if (d_in == nullptr || d_out == nullptr) return BFIR_ERR_INVALID;
    for (int b = 0; b < 8; b++) if (d_in[b] == nullptr || d_out[b] == nullptr) return BFIR_ERR_INVALID;
    if ((long long)e->impl.Ct * (e->impl.N / (e->impl.rs == 4 ? 4 : 2)) < 148LL * 128 * 4) {   // too small for the one-slice kernel
        rc = e->impl.shard_blocks_staged(4, d_in, d_out);
        return rc == BFIR_OK ? e->impl.shard_blocks_staged(4, d_in + 4, d_out + 4) : rc;
    }
    return e->impl.shard_blocks_staged(8, d_in, d_out);
}

int bfir_peer_setup(bfir_engine *e, int rank, int world)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.peer_setup(rank, world);
}

int bfir_peer_export(bfir_engine *e, void *handle64)
{
    if (e == nullptr || handle64 == nullptr || e->impl.recv == nullptr) return BFIR_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    BFIR_CUDA(cudaIpcGetMemHandle(&h, e->impl.recv));
    memcpy(handle64, &h, 64);
    return BFIR_OK;
}

int bfir_peer_import(bfir_engine *e, int peer_rank, const void *handle64)
{
    if (e == nullptr || handle64 == nullptr || !e->impl.peer.enabled || peer_rank < 0 || peer_rank >= e->impl.peer.world) return BFIR_ERR_INVALID;
    if (peer_rank == e->impl.peer.self) return BFIR_OK;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    BFIR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    e->impl.peer_opened[peer_rank] = p;
    e->impl.peer.recv[peer_rank] = p;
    return BFIR_OK;
}

int bfir_peer_set_ptr(bfir_engine *e, int peer_rank, void *d_recv)
{
    if (e == nullptr || d_recv == nullptr || !e->impl.peer.enabled || peer_rank < 0 || peer_rank >= e->impl.peer.world) return BFIR_ERR_INVALID;
    e->impl.peer.recv[peer_rank] = d_recv;
    return BFIR_OK;
}

void *bfir_peer_recv_ptr(bfir_engine *e) { return e ? e->impl.recv : nullptr; }

int bfir_peer_own_channels(bfir_engine *e, int *first, int *count)
{
    if (e == nullptr || !e->impl.peer.enabled) return BFIR_ERR_INVALID;
    if (first) *first = e->impl.own_first;
    if (count) *count = e->impl.own_count;
    return BFIR_OK;
}

void *bfir_acc_device_ptr(bfir_engine *e, size_t *bytes)
{
    if (e == nullptr) return nullptr;
    if (bytes) *bytes = (size_t)e->impl.N * e->impl.rs * e->impl.Ct;
    return e->impl.acc;
}

int bfir_sync(bfir_engine *e)
{
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.sync_and_probe();
}

int bfir_reset(bfir_engine *e)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.reset();
}

int bfir_get_overflow(bfir_engine *e, int channel, bfir_overflow_t *out)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.get_overflow(channel, out);
}

int bfir_check_overflows(bfir_engine *e)
{
    if (e != nullptr) e->impl.close_async();
    // brutefir::check_overflows + print_overflows, brutefir.cpp:371-388, 585-629
    if (e == nullptr) return BFIR_ERR_INVALID;
    Engine &g = e->impl;
    std::vector<bfir_overflow_t> cur(g.Cot);
    bool changed = false;
    for (int n = 0; n < g.Cot; n++) {
        int rc = g.get_overflow(n, &cur[n]);
        if (rc != BFIR_OK) return rc;
        if (memcmp(&cur[n], &g.last_overflow[n], sizeof(bfir_overflow_t)) != 0) changed = true;
    }
    if (!changed) return 0;
    g.last_overflow = cur;
    bool any = false;
    for (int n = 0; n < g.Cot; n++) if (cur[n].n_overflows > 0) { any = true; break; }
    if (!any) return 0;
    for (int n = 0; n < g.Cot; n++) {
        double peak = cur[n].largest;
        if (peak < (double)cur[n].intlargest) peak = (double)cur[n].intlargest;
        if (peak != 0.0) {
            if ((peak = 20.0 * log10(peak / cur[n].max)) == 0.0) peak = -0.0;
            bfir::pinfo("peak: %d/%u/%+.2f ", n, cur[n].n_overflows, peak);
        } else {
            bfir::pinfo("peak: %d/%u/-Inf ", n, cur[n].n_overflows);
        }
    }
    return 1;
}

int bfir_get_dither_ptr(bfir_engine *e, int channel, int *out)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr || out == nullptr || channel < 0 || channel >= e->impl.Cot) return BFIR_ERR_INVALID;
    if (!e->impl.dither_on) { set_error("engine has no dither state"); return BFIR_ERR_INVALID; }
    DitherState s;
    BFIR_CUDA(cudaStreamSynchronize(e->impl.stream));
    BFIR_CUDA(cudaMemcpy(&s, e->impl.dither.d_state + channel, sizeof(s), cudaMemcpyDeviceToHost));
    *out = s.randtab_ptr;
    return BFIR_OK;
}

int bfir_get_blockcounter(bfir_engine *e, unsigned int *out)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr || out == nullptr) return BFIR_ERR_INVALID;
    EngineState s;
    BFIR_CUDA(cudaStreamSynchronize(e->impl.stream));
    BFIR_CUDA(cudaMemcpy(&s, e->impl.state, sizeof(s), cudaMemcpyDeviceToHost));
    *out = s.blockcounter;
    return BFIR_OK;
}

int bfir_set_crossbar(bfir_engine *e, const double *in_gains, const double *out_gains)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.set_crossbar(in_gains, out_gains);
}

int bfir_set_groups(bfir_engine *e, int n_groups)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    return e->impl.set_groups(n_groups);
}

int bfir_get_groups(bfir_engine *e) { return e ? e->impl.n_groups : BFIR_ERR_INVALID; }
int bfir_get_mac_split(bfir_engine *e) { return e ? e->impl.mac_split : BFIR_ERR_INVALID; }
int bfir_get_quad_split(bfir_engine *e) { return e ? e->impl.quad_split : BFIR_ERR_INVALID; }

int bfir_set_stream(bfir_engine *e, void *cuda_stream)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr) return BFIR_ERR_INVALID;
    Engine &g = e->impl;
    // nothing queued on the previous stream (the engine's own or a caller's) may still be pending when work starts on
    // the new one: close_async() above ordered the side streams behind it, this drains it
    if (g.stream) BFIR_CUDA(cudaStreamSynchronize(g.stream));
    if (g.stream && g.own_stream) cudaStreamDestroy(g.stream);
    g.stream = (cudaStream_t)cuda_stream;
    g.own_stream = false;
    g.tail_ready = false;           // a look-ahead partition sum left on the old stream is not reused
    g.invalidate_graphs();
    return BFIR_OK;
}

int bfir_set_profiling(bfir_engine *e, int max_blocks)
{
    if (e != nullptr) e->impl.close_async();
    if (e == nullptr || max_blocks < 0) return BFIR_ERR_INVALID;
    Engine &g = e->impl;
    BFIR_CUDA(cudaStreamSynchronize(g.stream));
    g.prof_free();
    g.pev.resize((size_t)max_blocks * 4);
    for (auto &ev : g.pev) BFIR_CUDA(cudaEventCreate(&ev));
    g.pcap = (size_t)max_blocks;
    return BFIR_OK;
}

int bfir_get_mac_profile(bfir_engine *e, int blocks_per_launch, double *ms_sum, unsigned long long *launches, int reset)
{
    if (e == nullptr || ms_sum == nullptr || launches == nullptr || blocks_per_launch < 1 || blocks_per_launch > 8) return BFIR_ERR_INVALID;
    Engine &g = e->impl;
    *ms_sum = g.pmac_by_nb[blocks_per_launch];
    *launches = g.pcount_by_nb[blocks_per_launch];
    if (reset) { g.pmac_by_nb[blocks_per_launch] = 0.0; g.pcount_by_nb[blocks_per_launch] = 0; }
    return BFIR_OK;
}

int bfir_get_profile(bfir_engine *e, double ms_out[3], unsigned long long *blocks, int reset)
{
    if (e == nullptr || ms_out == nullptr || blocks == nullptr) return BFIR_ERR_INVALID;
    Engine &g = e->impl;
    for (int k = 0; k < 3; k++) ms_out[k] = g.pms[k];
    *blocks = g.pblocks;
    if (reset) { g.pms[0] = g.pms[1] = g.pms[2] = 0.0; g.pblocks = 0; }
    return BFIR_OK;
}

void bfir_set_print_callback(void (*cb)(const char *message)) { bfir::g_print_cb = cb; }
const char *bfir_last_error(void) { return bfir::g_last_error.c_str(); }
unsigned long long bfir_kernel_launch_count(void) { return bfir::g_launches.load(); }

}
