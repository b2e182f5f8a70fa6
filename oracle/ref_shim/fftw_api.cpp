// TEST INFRASTRUCTURE ONLY. The subset of the FFTW 3 API that the reference calls
// (brutefir/fftw_convolver.cpp:99-133, 204-209, 367-372, 384-397, 500-516, 556-561, 686-690, 798-806;
// brutefir/equalizer.cpp:262,357), provided on top of oracle/fft_r2r.hpp because neither libfftw3 nor
// libfftw3f exists on this host. Compiled against the reference's own header fftw/fftw3.h.
// Any number produced through this file must be labelled "FFT provider: oracle/fft_r2r (not FFTW)".
#include <fftw3.h>
#include "../fft_r2r.hpp"

namespace {
template <class T>
struct plan_impl {
    oracle_fft::RealFFT<T> fft;
    int kind;
    plan_impl(int n, int k) : fft(n), kind(k) {}
};
}

struct fftw_plan_s : plan_impl<double> { using plan_impl<double>::plan_impl; };
struct fftwf_plan_s : plan_impl<float> { using plan_impl<float>::plan_impl; };

extern "C" {

const char *oracle_fft_provider_name(void) { return "oracle/fft_r2r (own Stockham radix-4, native precision; NOT FFTW)"; }

fftw_plan fftw_plan_r2r_1d(int n, double *, double *, fftw_r2r_kind kind, unsigned)
{
    if (n < 1 || (n & (n - 1)) != 0 || (kind != FFTW_R2HC && kind != FFTW_HC2R)) return NULL;
    return new fftw_plan_s(n, (int)kind);
}

fftwf_plan fftwf_plan_r2r_1d(int n, float *, float *, fftwf_r2r_kind kind, unsigned)
{
    if (n < 1 || (n & (n - 1)) != 0 || (kind != FFTW_R2HC && kind != FFTW_HC2R)) return NULL;
    return new fftwf_plan_s(n, (int)kind);
}

void fftw_execute_r2r(const fftw_plan p, double *in, double *out)
{
    if (p->kind == FFTW_R2HC) p->fft.r2hc(in, out); else p->fft.hc2r(in, out);
}

void fftwf_execute_r2r(const fftwf_plan p, float *in, float *out)
{
    if (p->kind == FFTW_R2HC) p->fft.r2hc(in, out); else p->fft.hc2r(in, out);
}

void fftw_destroy_plan(fftw_plan p) { delete p; }
void fftwf_destroy_plan(fftwf_plan p) { delete p; }

// wisdom: nothing to learn, nothing to save
int fftw_import_wisdom(int (*)(void *), void *) { return 0; }
int fftwf_import_wisdom(int (*)(void *), void *) { return 0; }
void fftw_export_wisdom(void (*)(char, void *), void *) {}
void fftwf_export_wisdom(void (*)(char, void *), void *) {}

}
