// Device equalizer: counterpart of the reference's `class equalizer` (brutefir/equalizer.hpp:66-115,
// brutefir/equalizer.cpp) minus its WAV cache: generate() -> render -> the taps/2-sample filter stays
// on the device (or is copied to the host), ready for bfir_set_coeff_device.
#include "common.hpp"

namespace bfir {

static const double kIsoBands[31] = { 20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800,
    1000, 1250, 1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500, 16000, 20000 }; // equalizer.hpp:17-50

struct Equalizer {
    int block_length = 0, n_blocks = 0, rs = 0, rate = 0, taps = 0, log2m = 0, log2m1 = 0, log2m2 = 0;
    cudaStream_t stream = nullptr;
    void *tw = nullptr, *z = nullptr, *scratch = nullptr, *outbuf = nullptr; // z/scratch/outbuf: taps/2 complex each

    ~Equalizer() { destroy(); }
    int init(int bl, int nb, int realsize, int sampling_rate);
    void destroy();
    int render(int n_bands, const double *freq, const double *mag, const double *phase);
    const void *result() const { return (const char *)outbuf + (size_t)(taps / 2) * rs; } // upper half, equalizer.cpp:274-276
};

int Equalizer::init(int bl, int nb, int realsize, int sampling_rate)
{
    block_length = bl; n_blocks = nb; rs = realsize; rate = sampling_rate;
    if (rs != 4 && rs != 8) { set_error("Invalid real size %d.", rs); return BFIR_ERR_INVALID; }
    const long long total = (long long)bl * nb;
    const int lg = total > 0 && total <= (1 << 28) ? ilog2_exact((int)total) : -1;
    if (lg < 0) { set_error("Equalizer length (%d, %d) is not a power of two.", bl, nb); return BFIR_ERR_INVALID; } // equalizer.cpp:38-42
    taps = 1 << lg;
    log2m = lg - 1;                       // complex points of the packed real transform
    if (log2m < 4) { set_error("equalizer needs at least 32 taps"); return BFIR_ERR_INVALID; }
    const int maxsub = cfft_max_log2m(rs);
    if (log2m <= maxsub) { log2m1 = log2m; log2m2 = 0; }
    else { log2m1 = (log2m + 1) / 2; log2m2 = log2m - log2m1; }
    if (log2m1 > maxsub || (log2m2 != 0 && log2m2 < 4)) { set_error("equalizer length %d not supported", taps); return BFIR_ERR_INVALID; }
    BFIR_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc = make_twiddles(rs, taps, &tw);
    if (rc != BFIR_OK) return rc;
    const size_t bytes = (size_t)(taps / 2) * 2 * rs;
    BFIR_CUDA(cudaMalloc(&z, bytes));
    BFIR_CUDA(cudaMalloc(&scratch, bytes));
    BFIR_CUDA(cudaMalloc(&outbuf, bytes));
    return BFIR_OK;
}

void Equalizer::destroy()
{
    if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); stream = nullptr; }
    void *bufs[] = { tw, z, scratch, outbuf };
    for (void *b : bufs) if (b) cudaFree(b);
    tw = z = scratch = outbuf = nullptr;
}

int Equalizer::render(int n_bands, const double *freq, const double *mag, const double *phase)
{
    if (n_bands < 0 || n_bands > 31 || (n_bands > 0 && (freq == nullptr || mag == nullptr || phase == nullptr))) {
        set_error("Number of bands (%d) excceds limit (%d).", n_bands, 31);              // equalizer.cpp:96-100
        return BFIR_ERR_INVALID;
    }
    // equalizer::equalizer + generate, equalizer.cpp:57-66, 102-121. Unlike the reference instance (which
    // divides its band table by the rate in place on every call, :118) every render starts from clean tables.
    EqArgs a;
    memset(&a, 0, sizeof(a));
    const int bc = BFIR_EQ_BANDS;
    a.freq[0] = 0.0; a.freq[bc - 1] = (double)rate / 2.0;
    for (int n = 0; n < 31; n++) a.freq[n + 1] = kIsoBands[n];
    for (int n = 0, i = 0; n < n_bands; n++) {
        while (i < bc - 1 && freq[n] > a.freq[i]) i++;
        a.mag[i] = mag[n]; a.phase[i] = phase[n];
        i++;
        if (i >= bc) break;
    }
    a.mag[0] = a.mag[1]; a.mag[bc - 1] = a.mag[bc - 2];
    for (int n = 0; n < bc; n++) {
        a.freq[n] /= (double)rate;
        a.mag[n] = pow(10, a.mag[n] / 20);
        a.phase[n] /= (180 * M_PI);                                                       // sic, :120
    }
    a.zout = z; a.tw = tw; a.taps = taps;
    const int M = taps / 2, threads = 256;
    if (rs == 4) eq_spectrum_kernel<float><<<(M + threads - 1) / threads, threads, 0, stream>>>(a);
    else eq_spectrum_kernel<double><<<(M + threads - 1) / threads, threads, 0, stream>>>(a);
    count_launch();
    BFIR_CUDA(cudaGetLastError());
    // inverse complex transform of M points: z[n1 + M1 n2] = sum_k2 W_M2^(-k2 n2) W_M^(-k2 n1) sum_k1 Z[k1 M2 + k2] W_M1^(-k1 n1)
    const int M1 = 1 << log2m1, M2 = 1 << log2m2;
    CfftArgs c;
    memset(&c, 0, sizeof(c));
    c.tw = tw;
    if (log2m2 == 0) {
        c.in = z; c.out = outbuf; c.in_stride = 1; c.out_stride = 1; c.tw_shift_sub = 1; c.apply_tw = 0;
        cudaError_t e = launch_cfft_inverse(rs, log2m1, 1, stream, c);
        count_launch();
        if (e != cudaSuccess) { set_error("equalizer transform failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    } else {
        // step 1: M2 transforms of length M1 down the columns (stride M2), times W_M^(-k2 n1), in the same layout
        c.in = z; c.out = scratch; c.in_stride = M2; c.in_batch = 1; c.out_stride = M2; c.out_batch = 1;
        c.tw_shift_sub = 1 + log2m2;             // table length taps = 2 M = 2 M1 M2
        c.apply_tw = 1; c.tw_shift_tot = 1; c.mtot_mask = M - 1;
        cudaError_t e = launch_cfft_inverse(rs, log2m1, M2, stream, c);
        count_launch();
        if (e != cudaSuccess) { set_error("equalizer transform failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
        // step 2: M1 transforms of length M2 along the rows; result index n1 + M1 n2
        c.in = scratch; c.out = outbuf; c.in_stride = 1; c.in_batch = M2; c.out_stride = M1; c.out_batch = 1;
        c.tw_shift_sub = 1 + log2m1; c.apply_tw = 0;
        e = launch_cfft_inverse(rs, log2m2, M1, stream, c);
        count_launch();
        if (e != cudaSuccess) { set_error("equalizer transform failed: %s", cudaGetErrorString(e)); return BFIR_ERR_CUDA; }
    }
    return BFIR_OK;
}

} // namespace bfir

using namespace bfir;

struct bfir_eq { Equalizer impl; };

extern "C" {

int bfir_eq_create(bfir_eq **out, int block_length, int n_blocks, int realsize, int sampling_rate)
{
    if (out == nullptr) return BFIR_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_error("no CUDA device: libbfir_b200 has no CPU fallback"); return BFIR_ERR_CUDA; }
    bfir_eq *q = new bfir_eq;
    const int rc = q->impl.init(block_length, n_blocks, realsize, sampling_rate);
    if (rc != BFIR_OK) { delete q; return rc; }
    *out = q;
    return BFIR_OK;
}

void bfir_eq_destroy(bfir_eq *q) { delete q; }
int bfir_eq_taps(const bfir_eq *q) { return q ? q->impl.taps : BFIR_ERR_INVALID; }

int bfir_eq_render(bfir_eq *q, int n_bands, const double *freq, const double *mag, const double *phase, void *h_out)
{
    if (q == nullptr || h_out == nullptr) return BFIR_ERR_INVALID;
    const int rc = q->impl.render(n_bands, freq, mag, phase);
    if (rc != BFIR_OK) return rc;
    BFIR_CUDA(cudaMemcpyAsync(h_out, q->impl.result(), (size_t)(q->impl.taps / 2) * q->impl.rs, cudaMemcpyDeviceToHost, q->impl.stream));
    BFIR_CUDA(cudaStreamSynchronize(q->impl.stream));
    return BFIR_OK;
}

const void *bfir_eq_render_device(bfir_eq *q, int n_bands, const double *freq, const double *mag, const double *phase)
{
    if (q == nullptr) return nullptr;
    if (q->impl.render(n_bands, freq, mag, phase) != BFIR_OK) return nullptr;
    if (cudaStreamSynchronize(q->impl.stream) != cudaSuccess) return nullptr;
    return q->impl.result();
}

}
