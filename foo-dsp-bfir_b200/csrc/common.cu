// Shared host helpers: error text, print callback, twiddle tables, the dither tables.
#include "common.hpp"

namespace bfir {

thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches(0);
thread_local unsigned long long t_launches = 0;
void (*g_print_cb)(const char *) = nullptr;

void set_error(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

void pinfo(const char *fmt, ...)
{
    if (g_print_cb == nullptr) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_print_cb(buf);
}

int make_twiddles(int realsize, int n, void **d_out)
{
    *d_out = nullptr;
    const size_t bytes = (size_t)n * 2 * realsize;
    std::vector<unsigned char> h(bytes);
    for (int j = 0; j < n; j++) {
        const long double a = -2.0L * M_PIl * (long double)j / (long double)n;
        if (realsize == 4) { float *p = (float *)h.data(); p[2 * j] = (float)cosl(a); p[2 * j + 1] = (float)sinl(a); }
        else { double *p = (double *)h.data(); p[2 * j] = (double)cosl(a); p[2 * j + 1] = (double)sinl(a); }
    }
    BFIR_CUDA(cudaMalloc(d_out, bytes));
    BFIR_CUDA(cudaMemcpy(*d_out, h.data(), bytes, cudaMemcpyHostToDevice));
    return BFIR_OK;
}

// ---- dither.cpp:411-449: GSL "taus" maximally equidistributed combined Tausworthe generator
static inline uint32_t taus_step(uint32_t s, int a, int b, uint32_t c, int d) { return ((s & c) << d) ^ (((s << a) ^ s) >> b); }
static inline uint32_t tausrand(uint32_t st[3])
{
    st[0] = taus_step(st[0], 13, 19, 4294967294U, 12);
    st[1] = taus_step(st[1], 2, 25, 4294967288U, 4);
    st[2] = taus_step(st[2], 3, 11, 4294967280U, 17);
    return st[0] ^ st[1] ^ st[2];
}

int DitherTables::init(int n_ch, int sample_rate, int rs, int max_size, int max_samples_per_loop)
{
    n_channels = n_ch;
    realsize = rs;
    // table geometry, dither.cpp:29-60 (RANDTAB_SPACING 10 s, MIN_RANDTAB_SPACING 1 s)
    long long sp = 10LL * sample_rate;
    const long long minsp = (sample_rate > max_samples_per_loop) ? sample_rate : max_samples_per_loop;
    if (sp < minsp) sp = minsp;
    if (max_size > 0 && (long long)n_ch * sp > max_size) sp = max_size / n_ch;
    if (sp < minsp || (long long)n_ch * sp + 1 > 0x7fffffffLL) {
        set_error("dither table geometry invalid (channels %d, rate %d)", n_ch, sample_rate);
        return BFIR_ERR_INVALID;
    }
    spacing = (int)sp;
    size = n_ch * spacing + 1;
    uint32_t st[3];
    st[0] = 69069u * 1u;          // tausinit(state, 0): seed 0 -> 1, LCG chain (dither.cpp:429-440)
    st[1] = 69069u * st[0];
    st[2] = 69069u * st[1];
    for (int i = 0; i < 6; i++) tausrand(st); // warm-up, dither.cpp:442-448
    h_tab.resize(size);
    for (int n = 0; n < size; n++) h_tab[n] = (int8_t)(tausrand(st) & 0xFF);

    // TPDF map, dither.cpp:77-103. Index d = tab[n] - tab[n-1] in [-255, 255]; the reference table
    // stops at 254, entry 255 continues the formula (documented divergence: the reference reads
    // one element past its allocation there).
    h_map.assign(512, 0.0);
    std::vector<float> mf(512);
    h_map[0] = -0.5; mf[0] = -0.5f;
    for (int n = -255; n <= 255; n++) {
        volatile double term = 1.0 / 255.0 * (double)n;     // keep the reference's two roundings (no FMA)
        volatile double sum = 0.5 + 1.0 / 255.0;
        const double v = sum + term;
        h_map[n + 256] = v;
        mf[n + 256] = (float)v;
    }
    h_map[254 + 256] = 1.5; mf[254 + 256] = 1.5f;
    BFIR_CUDA(cudaMalloc((void **)&d_tab, (size_t)size));
    BFIR_CUDA(cudaMemcpy(d_tab, h_tab.data(), (size_t)size, cudaMemcpyHostToDevice));
    BFIR_CUDA(cudaMalloc(&d_map, 512 * (size_t)rs));
    if (rs == 4) {
        BFIR_CUDA(cudaMemcpy(d_map, mf.data(), 512 * sizeof(float), cudaMemcpyHostToDevice));
        for (int i = 0; i < 512; i++) h_map[i] = (double)mf[i];
    } else {
        BFIR_CUDA(cudaMemcpy(d_map, h_map.data(), 512 * sizeof(double), cudaMemcpyHostToDevice));
    }
    std::vector<DitherState> hs(n_ch);
    for (int n = 0; n < n_ch; n++) { // dither.cpp:105-109
        hs[n].randtab_ptr = n * spacing + 1;
        hs[n].tab0 = h_tab[0];
        hs[n].err[0] = hs[n].err[1] = 0.0;
    }
    BFIR_CUDA(cudaMalloc((void **)&d_state, sizeof(DitherState) * (size_t)n_ch));
    BFIR_CUDA(cudaMemcpy(d_state, hs.data(), sizeof(DitherState) * (size_t)n_ch, cudaMemcpyHostToDevice));
    return BFIR_OK;
}

void DitherTables::destroy()
{
    if (d_tab) cudaFree(d_tab);
    if (d_map) cudaFree(d_map);
    if (d_state) cudaFree(d_state);
    d_tab = nullptr; d_map = nullptr; d_state = nullptr;
}

} // namespace bfir

// pinned host memory for the raw blocks a caller hands to bfir_run (lets an integrator stay free of CUDA headers)
extern "C" void *bfir_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void bfir_host_free(void *p)
{
    if (p != nullptr) cudaFreeHost(p);
}
