// The convolver: device counterpart of the reference's `fftw_convolver` class
// (brutefir/fftw_convolver.hpp:28-166), one C entry point per public method, same argument order and
// the same T / HC / ORD buffer layouts, on device cbufs.
#include "common.hpp"

namespace bfir {

struct Conv {
    int L = 0, N = 0, rs = 0, log2m = 0, device = 0, n_dither = 0, fft_r0 = 1;
    cudaStream_t stream = nullptr;
    void *tw = nullptr;
    void *scratch = nullptr;       // one cbuf: in-place staging for two-CTA transforms, per instance
    void *stage = nullptr;         // half a cbuf: host coefficients on their way into coeffs2cbuf
    OverflowStats *d_stats = nullptr;
    int *d_flag = nullptr;
    DitherTables dither;
    bool own_stream = true;

    ~Conv() { destroy(); }
    // shared != nullptr: a helper instance (the td convolver's transforms) on its owner's stream, no dither tables
    int init(int length, int realsize, int n_dither_channels, int sampling_rate, const Conv *shared = nullptr);
    void destroy();
    // With two CTAs per transform an in-place call would let one CTA overwrite what the other still
    // has to read: use one CTA when the size allows, else go through the per-instance scratch cbuf.
    bool overlaps(const void *in, const void *out) const
    {
        const char *a = (const char *)in, *b = (const char *)out;
        const size_t n = (size_t)N * rs;
        return a < b + n && b < a + n;
    }
    int fwd(FwdArgs a)
    {
        int r0 = fft_r0;
        void *final_out = nullptr;
        if (r0 >= 2 && overlaps(a.in, a.out)) {   // several CTAs per transform cannot work in place
            if (rfft_choose_r0(rs, log2m, 1LL << 40) == 1) r0 = 1;
            else { final_out = a.out; a.out = scratch; }
        }
        cudaError_t e = launch_rfft_forward(rs, log2m, r0, dim3(1, 1), stream, a, tw);
        count_launch();
        if (e != cudaSuccess) { set_error("forward launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
        if (final_out) BFIR_CUDA(cudaMemcpyAsync(final_out, scratch, (size_t)N * rs, cudaMemcpyDeviceToDevice, stream));
        return BFIR_OK;
    }
    int inv(InvArgs a)
    {
        int r0 = fft_r0;
        void *final_out = nullptr;
        if (r0 >= 2 && overlaps(a.in, a.out)) {   // several CTAs per transform cannot work in place
            if (rfft_choose_r0(rs, log2m, 1LL << 40) == 1) r0 = 1;
            else { final_out = a.out; a.out = scratch; }
        }
        cudaError_t e = launch_rfft_inverse(rs, log2m, r0, dim3(1, 1), stream, a, tw);
        count_launch();
        if (e != cudaSuccess) { set_error("inverse launch failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
        if (final_out) BFIR_CUDA(cudaMemcpyAsync(final_out, scratch, (size_t)N * rs, cudaMemcpyDeviceToDevice, stream));
        return BFIR_OK;
    }
};

int Conv::init(int length, int realsize, int n_dither_channels, int sampling_rate, const Conv *shared)
{
    rs = realsize; L = length; N = 2 * length;
    if (rs != 4 && rs != 8) { set_error("Invalid real size %d.", rs); return BFIR_ERR_INVALID; }
    log2m = ilog2_exact(L);
    if (log2m < 0) { set_error("Invalid length %d.", L); return BFIR_ERR_INVALID; }
    if (!rfft_supported(rs, log2m)) { set_error("block length %d not supported for realsize %d", L, rs); return BFIR_ERR_INVALID; }
    fft_r0 = rfft_choose_r0(rs, log2m, 1);
    BFIR_CUDA(cudaGetDevice(&device));
    if (shared) { stream = shared->stream; own_stream = false; }
    else BFIR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc = make_twiddles(rs, N, &tw);
    if (rc != BFIR_OK) return rc;
    BFIR_CUDA(cudaMalloc(&scratch, (size_t)N * rs));
    BFIR_CUDA(cudaMalloc(&stage, (size_t)L * rs));
    BFIR_CUDA(cudaMalloc((void **)&d_stats, sizeof(OverflowStats)));
    BFIR_CUDA(cudaMalloc((void **)&d_flag, sizeof(int)));
    if (shared) return BFIR_OK;
    n_dither = n_dither_channels > 0 ? n_dither_channels : 1;
    return dither.init(n_dither, sampling_rate, rs, 0, L);
}

void Conv::destroy()
{
    if (stream) { cudaStreamSynchronize(stream); if (own_stream) cudaStreamDestroy(stream); stream = nullptr; }
    if (tw) cudaFree(tw);
    if (scratch) cudaFree(scratch);
    if (stage) cudaFree(stage);
    if (d_stats) cudaFree(d_stats);
    if (d_flag) cudaFree(d_flag);
    tw = scratch = stage = nullptr; d_stats = nullptr; d_flag = nullptr;
    dither.destroy();
}

} // namespace bfir

using namespace bfir;

struct bfir_conv { Conv impl; };
struct bfir_td_conv {
    Conv fft;                 // transforms of 2 * blocklen points on the owner's stream
    void *d_coeffs = nullptr; // spectrum of [0 | h | 0] / (2 blocklen), plain HC layout
    int blocklen = 0;
};

#define CONV_CHECK(c) do { if ((c) == nullptr) return BFIR_ERR_INVALID; } while (0)
#define LAUNCH_1D(kernel, n, ...)                                                   \
    do {                                                                            \
        const int threads__ = 256, blocks__ = ((n) + threads__ - 1) / threads__;    \
        kernel<<<blocks__, threads__, 0, g.stream>>>(__VA_ARGS__);                  \
        count_launch();                                                             \
        BFIR_CUDA(cudaGetLastError());                                              \
    } while (0)

extern "C" {

int bfir_conv_create(bfir_conv **out, int length, int realsize, int n_dither_channels, int sampling_rate)
{
    if (out == nullptr) return BFIR_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        set_error("no CUDA device: libbfir_b200 has no CPU fallback");
        return BFIR_ERR_CUDA;
    }
    bfir_conv *c = new bfir_conv;
    const int rc = c->impl.init(length, realsize, n_dither_channels, sampling_rate);
    if (rc != BFIR_OK) { delete c; return rc; }
    *out = c;
    return BFIR_OK;
}

void bfir_conv_destroy(bfir_conv *c) { delete c; }
int bfir_conv_cbufsize(const bfir_conv *c) { return c ? c->impl.N * c->impl.rs : BFIR_ERR_INVALID; }

void *bfir_conv_alloc(bfir_conv *c, size_t bytes)
{
    if (c == nullptr || bytes == 0) return nullptr;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    cudaMemsetAsync(p, 0, bytes, c->impl.stream);
    return p;
}

void bfir_conv_free(bfir_conv *c, void *d_ptr)
{
    if (c != nullptr) cudaStreamSynchronize(c->impl.stream);
    if (d_ptr) cudaFree(d_ptr);
}

int bfir_conv_upload(bfir_conv *c, void *d_dst, const void *h_src, size_t bytes)
{
    CONV_CHECK(c);
    BFIR_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->impl.stream));
    BFIR_CUDA(cudaStreamSynchronize(c->impl.stream));
    return BFIR_OK;
}

int bfir_conv_download(bfir_conv *c, void *h_dst, const void *d_src, size_t bytes)
{
    CONV_CHECK(c);
    BFIR_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->impl.stream));
    BFIR_CUDA(cudaStreamSynchronize(c->impl.stream));
    return BFIR_OK;
}

int bfir_conv_sync(bfir_conv *c)
{
    CONV_CHECK(c);
    BFIR_CUDA(cudaStreamSynchronize(c->impl.stream));
    return BFIR_OK;
}

int bfir_conv_raw2cbuf(bfir_conv *c, const void *d_rawbuf, void *cbuf, void *next_cbuf, int format, int byte_offset, int sample_spacing)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (!fmt_valid(format) || sample_spacing < 1) return BFIR_ERR_INVALID;
    const uint8_t *raw = (const uint8_t *)d_rawbuf + byte_offset;
    if (g.rs == 4) LAUNCH_1D(raw2cbuf_kernel<float>, g.L, raw, (float *)cbuf, (float *)next_cbuf, format, sample_spacing, g.L);
    else LAUNCH_1D(raw2cbuf_kernel<double>, g.L, raw, (double *)cbuf, (double *)next_cbuf, format, sample_spacing, g.L);
    return BFIR_OK;
}

int bfir_conv_time2freq(bfir_conv *c, const void *input_cbuf, void *output_cbuf)
{
    CONV_CHECK(c);
    FwdArgs a = {};
    a.in_mode = IN_TIME; a.out_layout = LAYOUT_HC; a.in = input_cbuf; a.out = output_cbuf; a.scale_in = 1.0; a.scale_out = 1.0;
    return c->impl.fwd(a);
}

int bfir_conv_freq2time(bfir_conv *c, const void *input_cbuf, void *output_cbuf)
{
    CONV_CHECK(c);
    InvArgs a = {};
    a.in_layout = LAYOUT_HC; a.out_mode = OUT_TIME; a.in = input_cbuf; a.out = output_cbuf; a.scale_in = 1.0;
    return c->impl.inv(a);
}

int bfir_conv_mixnscale(bfir_conv *c, void *const *input_cbufs, void *output_cbuf, const double *scales, int n_bufs, int mixmode)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (mixmode != MIXMODE_INPUT && mixmode != MIXMODE_OUTPUT) { // fftw_convolver.cpp:1423-1425
        pinfo("Invalid mixmode: %d.\n", mixmode);
        set_error("Invalid mixmode: %d.", mixmode);
        return BFIR_ERR_INVALID;
    }
    if (n_bufs < 1 || n_bufs > BFIR_MAX_MIX_BUFS || input_cbufs == nullptr || scales == nullptr) return BFIR_ERR_INVALID;
    MixArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < n_bufs; i++) {
        if (input_cbufs[i] == output_cbuf) { set_error("mixnscale: output aliases input %d", i); return BFIR_ERR_INVALID; }
        a.in[i] = input_cbufs[i];
        // the float build narrows the scales first (fftw_convolver.cpp:871-876)
        a.scales[i] = g.rs == 4 ? (double)(float)scales[i] : scales[i];
    }
    a.out = output_cbuf; a.n_bufs = n_bufs; a.mixmode = mixmode; a.N = g.N;
    if (g.rs == 4) LAUNCH_1D(mixnscale_kernel<float>, g.N, a);
    else LAUNCH_1D(mixnscale_kernel<double>, g.N, a);
    return BFIR_OK;
}

int bfir_conv_convolve(bfir_conv *c, const void *input_cbuf, const void *coeffs, void *output_cbuf)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (g.rs == 4) LAUNCH_1D((convolve_kernel<float, 0>), g.N / 8, (const float *)input_cbuf, (const float *)coeffs, (float *)output_cbuf, g.N);
    else LAUNCH_1D((convolve_kernel<double, 0>), g.N / 8, (const double *)input_cbuf, (const double *)coeffs, (double *)output_cbuf, g.N);
    return BFIR_OK;
}

int bfir_conv_convolve_inplace(bfir_conv *c, void *cbuf, const void *coeffs)
{
    // fftw_convolver.cpp:1430-1462: every group of 8 is read before it is written, so out == in is safe
    return bfir_conv_convolve(c, cbuf, coeffs, cbuf);
}

int bfir_conv_convolve_add(bfir_conv *c, const void *input_cbuf, const void *coeffs, void *output_cbuf)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (g.rs == 4) LAUNCH_1D((convolve_kernel<float, 1>), g.N / 8, (const float *)input_cbuf, (const float *)coeffs, (float *)output_cbuf, g.N);
    else LAUNCH_1D((convolve_kernel<double, 1>), g.N / 8, (const double *)input_cbuf, (const double *)coeffs, (double *)output_cbuf, g.N);
    return BFIR_OK;
}

int bfir_conv_crossfade_inplace(bfir_conv *c, void *input_cbuf, void *crossfade_cbuf, void *buffer_cbuf)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    // (1) crossfade = HC2R(ORD->HC(crossfade)): old filter's time-domain output   fftw_convolver.cpp:289-291
    InvArgs a = {};
    a.in_layout = LAYOUT_ORD; a.out_mode = OUT_TIME; a.scale_in = 1.0;
    a.in = crossfade_cbuf; a.out = crossfade_cbuf;
    int rc = g.inv(a);
    if (rc != BFIR_OK) return rc;
    // (2) buffer = HC2R(ORD->HC(input)): new filter's time-domain output          :292-294
    a.in = input_cbuf; a.out = buffer_cbuf;
    rc = g.inv(a);
    if (rc != BFIR_OK) return rc;
    // (3) linear old->new ramp over the valid first half                          :296-305
    if (g.rs == 4) LAUNCH_1D(crossfade_ramp_kernel<float>, g.L, (const float *)crossfade_cbuf, (float *)buffer_cbuf, g.L);
    else LAUNCH_1D(crossfade_ramp_kernel<double>, g.L, (const double *)crossfade_cbuf, (double *)buffer_cbuf, g.L);
    // (4)+(5) input = HC->ORD(R2HC(buffer)) / N                                   :317-320
    FwdArgs f = {};
    f.in_mode = IN_TIME; f.out_layout = LAYOUT_ORD; f.in = buffer_cbuf; f.out = input_cbuf; f.scale_in = 1.0; f.scale_out = 1.0 / (double)g.N;
    return g.fwd(f);
}

int bfir_conv_dirac_convolve(bfir_conv *c, const void *input_cbuf, void *output_cbuf)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (g.rs == 4) LAUNCH_1D(dirac_kernel<float>, g.N, (const float *)input_cbuf, (float *)output_cbuf, g.N);
    else LAUNCH_1D(dirac_kernel<double>, g.N, (const double *)input_cbuf, (double *)output_cbuf, g.N);
    return BFIR_OK;
}

int bfir_conv_dirac_convolve_inplace(bfir_conv *c, void *cbuf) { return bfir_conv_dirac_convolve(c, cbuf, cbuf); }

int bfir_conv_convolve_eval(bfir_conv *c, const void *input_cbuf, void *buffer_cbuf, void *output_cbuf)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    // buffer[L .. L+N) = HC2R(input)                                              fftw_convolver.cpp:384-386
    InvArgs a = {};
    a.in_layout = LAYOUT_HC; a.out_mode = OUT_TIME; a.scale_in = 1.0;
    a.in = input_cbuf; a.out = (char *)buffer_cbuf + (size_t)g.L * g.rs;
    int rc = g.inv(a);
    if (rc != BFIR_OK) return rc;
    // output = R2HC(buffer[0 .. N))                                                :387-389
    FwdArgs f = {};
    f.in_mode = IN_TIME; f.out_layout = LAYOUT_HC; f.in = buffer_cbuf; f.out = output_cbuf; f.scale_in = 1.0; f.scale_out = 1.0;
    rc = g.fwd(f);
    if (rc != BFIR_OK) return rc;
    // buffer[0 .. L) = buffer[L .. 2L)                                             :401-402
    BFIR_CUDA(cudaMemcpyAsync(buffer_cbuf, (char *)buffer_cbuf + (size_t)g.L * g.rs, (size_t)g.L * g.rs, cudaMemcpyDeviceToDevice, g.stream));
    return BFIR_OK;
}

// log2_roof of the reference (log2.h:34-51) for n >= 2; n == 1 shifts by -1 there and is refused here
int bfir_conv_td_block_length(int n_coeffs)
{
    if (n_coeffs < 2 || n_coeffs > (1 << 30)) return -1;
    int lg = 0;
    while ((1 << lg) < n_coeffs) lg++;
    return 1 << lg;
}

int bfir_conv_td_new(bfir_conv *c, bfir_td_conv **out, const void *h_coeffs, int n_coeffs)
{
    CONV_CHECK(c);
    if (out == nullptr || h_coeffs == nullptr) return BFIR_ERR_INVALID;
    *out = nullptr;
    Conv &g = c->impl;
    const int bl = bfir_conv_td_block_length(n_coeffs);
    if (bl < 0) { set_error("td_new: invalid coefficient count %d", n_coeffs); return BFIR_ERR_INVALID; }
    bfir_td_conv *t = new bfir_td_conv;
    int rc = t->fft.init(bl, g.rs, 0, 0, &g);
    if (rc != BFIR_OK) { delete t; return rc; }
    t->blocklen = bl;
    const size_t rs = (size_t)g.rs;
    if (cudaMalloc(&t->d_coeffs, 2 * (size_t)bl * rs) != cudaSuccess) { delete t; set_error("td_new: out of device memory"); return BFIR_ERR_CUDA; }
    // [0_blocklen | coeffs | 0]                                                     fftw_convolver.cpp:731-736
    cudaError_t e = cudaMemsetAsync(t->d_coeffs, 0, 2 * (size_t)bl * rs, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync((char *)t->d_coeffs + (size_t)bl * rs, h_coeffs, (size_t)n_coeffs * rs, cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream); // h_coeffs is the caller's, pageable
    if (e != cudaSuccess) { bfir_conv_td_free(t); set_error("td_new: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    // in-place R2HC, then every element times 1 / (2 blocklen)                       :738-757
    FwdArgs a = {};
    a.in_mode = IN_TIME; a.out_layout = LAYOUT_HC; a.in = t->d_coeffs; a.out = t->d_coeffs;
    a.scale_in = 1.0; a.scale_out = 1.0 / (double)(bl << 1);
    rc = t->fft.fwd(a);
    if (rc != BFIR_OK) { bfir_conv_td_free(t); return rc; }
    *out = t;
    return BFIR_OK;
}

int bfir_conv_td_blocklen(const bfir_td_conv *tdc) { return tdc ? tdc->blocklen : BFIR_ERR_INVALID; }
const void *bfir_conv_td_coeffs(const bfir_td_conv *tdc) { return tdc ? tdc->d_coeffs : nullptr; }

int bfir_conv_td_convolve(bfir_conv *c, bfir_td_conv *tdc, void *d_overlap_block)
{
    CONV_CHECK(c);
    if (tdc == nullptr || d_overlap_block == nullptr) return BFIR_ERR_INVALID;
    Conv &g = c->impl;
    if (tdc->fft.stream != g.stream || tdc->fft.rs != g.rs) { set_error("td_convolve: td_conv_t belongs to another convolver"); return BFIR_ERR_INVALID; }
    const int size = tdc->blocklen << 1;
    FwdArgs f = {};                                                                // :763-777
    f.in_mode = IN_TIME; f.out_layout = LAYOUT_HC; f.in = d_overlap_block; f.out = d_overlap_block; f.scale_in = 1.0; f.scale_out = 1.0;
    int rc = tdc->fft.fwd(f);
    if (rc != BFIR_OK) return rc;
    if (g.rs == 4) LAUNCH_1D(hc_convolve_inplace_kernel<float>, size / 2 + 1, (float *)d_overlap_block, (const float *)tdc->d_coeffs, size);
    else LAUNCH_1D(hc_convolve_inplace_kernel<double>, size / 2 + 1, (double *)d_overlap_block, (const double *)tdc->d_coeffs, size);
    InvArgs a = {};
    a.in_layout = LAYOUT_HC; a.out_mode = OUT_TIME; a.in = d_overlap_block; a.out = d_overlap_block; a.scale_in = 1.0;
    return tdc->fft.inv(a);
}

void bfir_conv_td_free(bfir_td_conv *tdc)
{
    if (tdc == nullptr) return;
    if (tdc->fft.stream) cudaStreamSynchronize(tdc->fft.stream);
    if (tdc->d_coeffs) cudaFree(tdc->d_coeffs);
    delete tdc;
}

int bfir_conv_cbuf2raw(bfir_conv *c, const void *cbuf, void *d_outbuf, int format, int byte_offset, int sample_spacing,
                       int apply_dither, int dither_channel, bfir_overflow_t *overflow)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (!fmt_valid(format) || sample_spacing < 1 || overflow == nullptr) return BFIR_ERR_INVALID;
    if (dither_channel < 0 || dither_channel >= g.n_dither) return BFIR_ERR_INVALID;
    OverflowStats s;
    s.n_overflows = overflow->n_overflows; s.intlargest = overflow->intlargest;
    memcpy(&s.largest_bits, &overflow->largest, sizeof(double));
    BFIR_CUDA(cudaMemcpyAsync(g.d_stats, &s, sizeof(s), cudaMemcpyHostToDevice, g.stream));
    uint8_t *raw = (uint8_t *)d_outbuf + byte_offset;
    if (apply_dither && !fmt_isfloat(format)) {                                      // fftw_convolver.cpp:421,444
        DitherArgs d = {};
        d.real = cbuf; d.real_stride = 0; d.raw = raw; d.raw_stream_stride = 0;
        d.fmt = format; d.ch_per_stream = sample_spacing; d.L = g.L; d.n_channels = 1;
        d.randtab = g.dither.d_tab; d.randtab_size = g.dither.size; d.randmap = g.dither.d_map;
        d.dstate = g.dither.d_state; d.stats = g.d_stats - dither_channel; d.single_channel = dither_channel;
        if (g.rs == 4) dither_kernel<float><<<1, 128, 0, g.stream>>>(d, 1);
        else dither_kernel<double><<<1, 128, 0, g.stream>>>(d, 1);
        count_launch();
        BFIR_CUDA(cudaGetLastError());
    } else {
        if (g.rs == 4) LAUNCH_1D(real2raw_kernel<float>, g.L, (const float *)cbuf, raw, format, sample_spacing, g.L, overflow->max, g.d_stats);
        else LAUNCH_1D(real2raw_kernel<double>, g.L, (const double *)cbuf, raw, format, sample_spacing, g.L, overflow->max, g.d_stats);
    }
    BFIR_CUDA(cudaMemcpyAsync(&s, g.d_stats, sizeof(s), cudaMemcpyDeviceToHost, g.stream));
    BFIR_CUDA(cudaStreamSynchronize(g.stream));
    overflow->n_overflows = s.n_overflows; overflow->intlargest = s.intlargest;
    memcpy(&overflow->largest, &s.largest_bits, sizeof(double));
    return BFIR_OK;
}

int bfir_conv_coeffs2cbuf(bfir_conv *c, const void *coeffs, int n_coeffs, double scale, void *d_dest)
{
    CONV_CHECK(c);
    Conv &g = c->impl;
    if (coeffs == nullptr || d_dest == nullptr || n_coeffs < 0) return BFIR_ERR_INVALID;
    const int len = n_coeffs > g.L ? g.L : n_coeffs;                                 // fftw_convolver.cpp:483
    if (len > 0) BFIR_CUDA(cudaMemcpyAsync(g.stage, coeffs, (size_t)len * g.rs, cudaMemcpyHostToDevice, g.stream));
    BFIR_CUDA(cudaMemsetAsync(g.d_flag, 0, sizeof(int), g.stream));
    FwdArgs a = {};
    a.in_mode = IN_COEFF; a.out_layout = LAYOUT_ORD; a.in = g.stage; a.out = d_dest;
    a.scale_in = scale; a.scale_out = 1.0 / (double)g.N; a.coeff_len = len; a.nonfinite = g.d_flag;
    int rc = g.fwd(a);
    if (rc != BFIR_OK) return rc;
    int bad = 0;
    BFIR_CUDA(cudaMemcpyAsync(&bad, g.d_flag, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
    BFIR_CUDA(cudaStreamSynchronize(g.stream));
    if (bad) { pinfo("NaN or Inf value among coefficients.\n"); set_error("NaN or Inf value among coefficients."); return BFIR_ERR_COEFF; }
    return BFIR_OK;
}

int bfir_conv_runtime_coeffs2cbuf(bfir_conv *c, const void *d_src, void *d_dest)
{
    CONV_CHECK(c);
    // dest = HC->ORD(R2HC([0 | src])) / N, fftw_convolver.cpp:551-566. All loads precede all stores
    // inside the CTA, so src may live in the upper half of dest like in the reference's callers.
    FwdArgs a = {};
    a.in_mode = IN_UPPER; a.out_layout = LAYOUT_ORD; a.in = d_src; a.out = d_dest; a.scale_in = 1.0; a.scale_out = 1.0 / (double)c->impl.N;
    return c->impl.fwd(a);
}

int bfir_conv_dither_table_size(bfir_conv *c) { return c ? c->impl.dither.size : BFIR_ERR_INVALID; }

int bfir_conv_dither_table(bfir_conv *c, int8_t *h_out, int n)
{
    CONV_CHECK(c);
    if (h_out == nullptr || n < 0 || n > c->impl.dither.size) return BFIR_ERR_INVALID;
    BFIR_CUDA(cudaStreamSynchronize(c->impl.stream));
    BFIR_CUDA(cudaMemcpy(h_out, c->impl.dither.d_tab, (size_t)n, cudaMemcpyDeviceToHost)); // what the kernels read
    return BFIR_OK;
}

int bfir_conv_dither_map(bfir_conv *c, void *h_out)
{
    CONV_CHECK(c);
    BFIR_CUDA(cudaMemcpy(h_out, c->impl.dither.d_map, 512 * (size_t)c->impl.rs, cudaMemcpyDeviceToHost));
    return BFIR_OK;
}

int bfir_conv_dither_ptr(bfir_conv *c, int channel)
{
    if (c == nullptr || channel < 0 || channel >= c->impl.n_dither) return BFIR_ERR_INVALID;
    DitherState s;
    if (cudaStreamSynchronize(c->impl.stream) != cudaSuccess) return BFIR_ERR_CUDA;
    if (cudaMemcpy(&s, c->impl.dither.d_state + channel, sizeof(s), cudaMemcpyDeviceToHost) != cudaSuccess) return BFIR_ERR_CUDA;
    return s.randtab_ptr;
}

}
