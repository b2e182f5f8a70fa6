// TEST INFRASTRUCTURE ONLY. Second provider of the FFTW 3 calls the reference makes (see fftw_api.cpp for the
// list): Intel MKL's DFTI interface as exported by PyTorch's libtorch_cpu.so -- the tuned CPU FFT this image has
// (provider 2; libfftw3 itself is not installed). Used for the CPU baseline timings only
// (bench.py cpu_baseline / --impl reference); the bit-exact parity tests stay on oracle/fft_r2r.
//
// FFTW_R2HC / FFTW_HC2R (power-of-two n, unnormalised) map onto a real 1-D DFTI descriptor with
// conjugate-even storage as n/2+1 complex values (CCE), out of place into a per-plan scratch, plus one repacking
// pass between CCE and FFTW's half-complex order. Any number produced through this file is labelled
// "FFT provider: MKL DFTI (libtorch_cpu.so) + half-complex repack (NOT FFTW)".
#include <fftw3.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <complex>
#include <mutex>
#include <vector>

namespace {

// the handful of DFTI names and constants used here (Intel MKL mkl_dfti.h; the header is not in this image)
typedef void *dfti_handle;
enum { DFTI_PLACEMENT = 11, DFTI_CONJUGATE_EVEN_STORAGE = 10, DFTI_THREAD_LIMIT = 27, DFTI_NUMBER_OF_USER_THREADS = 26 };
enum { DFTI_REAL = 33, DFTI_COMPLEX_COMPLEX = 39, DFTI_NOT_INPLACE = 44 };
typedef long (*create_1d_t)(dfti_handle *, int domain, long length);
typedef long (*set_value_t)(dfti_handle, int param, ...);
typedef long (*commit_t)(dfti_handle);
typedef long (*compute_t)(dfti_handle, void *, ...);
typedef long (*free_t)(dfti_handle *);

struct Mkl {
    create_1d_t create_d = nullptr, create_s = nullptr;
    set_value_t set_value = nullptr;
    commit_t commit = nullptr;
    compute_t forward = nullptr, backward = nullptr;
    free_t release = nullptr;
    bool ok = false;
};

Mkl &mkl()
{
    static Mkl m;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *path = getenv("BFIR_MKL_LIB");
        void *h = dlopen(path && *path ? path : "libtorch_cpu.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        m.create_d = (create_1d_t)dlsym(h, "DftiCreateDescriptor_d_1d");
        m.create_s = (create_1d_t)dlsym(h, "DftiCreateDescriptor_s_1d");
        m.set_value = (set_value_t)dlsym(h, "DftiSetValue");
        m.commit = (commit_t)dlsym(h, "DftiCommitDescriptor");
        m.forward = (compute_t)dlsym(h, "DftiComputeForward");
        m.backward = (compute_t)dlsym(h, "DftiComputeBackward");
        m.release = (free_t)dlsym(h, "DftiFreeDescriptor");
        m.ok = m.create_d && m.create_s && m.set_value && m.commit && m.forward && m.backward && m.release;
    });
    return m;
}

template <class T>
struct plan_impl {
    dfti_handle desc = nullptr;
    int n, kind;
    std::vector<std::complex<T>> cce;   // n/2 + 1 values
    plan_impl(int n_, int k) : n(n_), kind(k), cce((size_t)n_ / 2 + 1) {}
    bool init()
    {
        Mkl &m = mkl();
        if (!m.ok) return false;
        if ((sizeof(T) == 8 ? m.create_d : m.create_s)(&desc, DFTI_REAL, (long)n) != 0) return false;
        if (m.set_value(desc, DFTI_PLACEMENT, DFTI_NOT_INPLACE) != 0) return false;
        if (m.set_value(desc, DFTI_CONJUGATE_EVEN_STORAGE, DFTI_COMPLEX_COMPLEX) != 0) return false;
        m.set_value(desc, DFTI_THREAD_LIMIT, 1L);   // one engine per host thread, like the reference
        return m.commit(desc) == 0;
    }
    ~plan_impl() { if (desc) mkl().release(&desc); }
    // hc[k] = Re X_k (0 <= k <= n/2), hc[n-k] = Im X_k (0 < k < n/2)
    void r2hc(T *in, T *out)
    {
        mkl().forward(desc, in, cce.data());
        const int h = n / 2;
        out[0] = cce[0].real();
        for (int k = 1; k < h; k++) { out[k] = cce[k].real(); out[n - k] = cce[k].imag(); }
        out[h] = cce[h].real();
    }
    void hc2r(T *in, T *out)
    {
        const int h = n / 2;
        cce[0] = std::complex<T>(in[0], 0);
        for (int k = 1; k < h; k++) cce[k] = std::complex<T>(in[k], in[n - k]);
        cce[h] = std::complex<T>(in[h], 0);
        mkl().backward(desc, cce.data(), out);
    }
};
}

struct fftw_plan_s : plan_impl<double> { using plan_impl<double>::plan_impl; };
struct fftwf_plan_s : plan_impl<float> { using plan_impl<float>::plan_impl; };

extern "C" {

const char *oracle_fft_provider_name(void)
{
    return mkl().ok ? "MKL DFTI (libtorch_cpu.so) + half-complex repack (NOT FFTW)" : NULL;
}

fftw_plan fftw_plan_r2r_1d(int n, double *, double *, fftw_r2r_kind kind, unsigned)
{
    if (n < 2 || (n & (n - 1)) != 0 || (kind != FFTW_R2HC && kind != FFTW_HC2R)) return NULL;
    fftw_plan p = new fftw_plan_s(n, (int)kind);
    if (!p->init()) { delete p; return NULL; }
    return p;
}

fftwf_plan fftwf_plan_r2r_1d(int n, float *, float *, fftwf_r2r_kind kind, unsigned)
{
    if (n < 2 || (n & (n - 1)) != 0 || (kind != FFTW_R2HC && kind != FFTW_HC2R)) return NULL;
    fftwf_plan p = new fftwf_plan_s(n, (int)kind);
    if (!p->init()) { delete p; return NULL; }
    return p;
}

void fftw_execute_r2r(const fftw_plan p, double *in, double *out)
{
    if (p->kind == FFTW_R2HC) p->r2hc(in, out); else p->hc2r(in, out);
}

void fftwf_execute_r2r(const fftwf_plan p, float *in, float *out)
{
    if (p->kind == FFTW_R2HC) p->r2hc(in, out); else p->hc2r(in, out);
}

void fftw_destroy_plan(fftw_plan p) { delete p; }
void fftwf_destroy_plan(fftwf_plan p) { delete p; }

int fftw_import_wisdom(int (*)(void *), void *) { return 0; }
int fftwf_import_wisdom(int (*)(void *), void *) { return 0; }
void fftw_export_wisdom(void (*)(char, void *), void *) {}
void fftwf_export_wisdom(void (*)(char, void *), void *) {}

}
