// N3 parity: host/preprocessor.hpp (the offline drivers on the GPU engine) against the reference's OWN
// brutefir/preprocessor.cpp, compiled unmodified into oracle/_ref/libbfir_ref.so and fed the same dense impulse
// responses through in-memory "sound files" (oracle/ref_shim/stubs.cpp). TEST ONLY: links the oracle.
//   convolve_impulses      (preprocessor.cpp:34-233)   cascade of three responses, per-file scales
//   calculate_attenuation  (preprocessor.cpp:250-412)  same white noise on both sides
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include "../../foo-dsp-bfir_b200/host/preprocessor.hpp"

extern "C" {
void ref_vfile_register(const wchar_t *name, int channels, int frames, int rate, const double *interleaved);
void ref_vfile_clear(void);
void ref_set_noise(const double *samples, long long n);
int ref_convolve_impulses(const wchar_t *const *names, const double *scales, int n, int filter_length, int realsize, void *out,
                          long long capacity, int *channels);
int ref_calculate_attenuation(const wchar_t *name, int filter_length, int realsize, double *attenuation, int *n_channels,
                              int *n_frames, int *sampling_rate);
}

template <class T> static int run(const char *tag, double tol, double tol_db)
{
    const int L = 256, C = 3, rate = 48000;
    const int frames[3] = { 700, 1000, 333 };
    const wchar_t *names[3] = { L"a.wav", L"b.wav", L"c.wav" };
    const double scales[3] = { 1.0, 0.5, 2.0 };
    std::mt19937 gen(7);
    std::normal_distribution<double> nd(0.0, 1.0);
    std::vector<std::vector<std::vector<T> > > imps(3, std::vector<std::vector<T> >(C));
    ref_vfile_clear();
    for (int k = 0; k < 3; k++) {
        std::vector<double> inter((size_t)frames[k] * C);
        for (int f = 0; f < frames[k]; f++)
            for (int c = 0; c < C; c++) {
                // dense decaying response, rounded to T so that both sides start from identical samples
                const T v = (T)(nd(gen) * std::exp(-3.0 * f / frames[k]) * 0.2);
                inter[(size_t)f * C + c] = (double)v;
            }
        for (int c = 0; c < C; c++) {
            imps[k][c].resize(frames[k]);
            for (int f = 0; f < frames[k]; f++) imps[k][c][f] = (T)inter[(size_t)f * C + c];
        }
        ref_vfile_register(names[k], C, frames[k], rate, inter.data());
    }
    const int P = (1000 + L - 1) / L;                               // util::get_next_multiple(g_frames, L) / L
    std::vector<T> ref((size_t)L * P * C);
    int ch = 0;
    const int n = ref_convolve_impulses(names, scales, 3, L, (int)sizeof(T), ref.data(), (long long)(ref.size() * sizeof(T)), &ch);
    if (n != 1000 || ch != C) { printf("%s: reference cascade failed (%d frames, %d channels)\n", tag, n, ch); return 1; }
    std::vector<std::vector<T> > got;
    if (!preprocessor::convolve_impulses<T>(imps, std::vector<double>(scales, scales + 3), L, P, rate, got)) { printf("%s: cascade failed\n", tag); return 1; }
    double num = 0, den = 0;
    for (int f = 0; f < n; f++)
        for (int c = 0; c < C; c++) {
            const double a = got[c][f], b = ref[(size_t)f * C + c];
            num += (a - b) * (a - b); den += b * b;
        }
    const double err = std::sqrt(num / den);
    printf("%s: convolve_impulses rel rms vs reference %.3e (%d frames, energy %.3e)\n", tag, err, n, den);
    if (!(err < tol) || !(den > 0)) return 1;

    // calculate_attenuation on the loudest response, identical noise
    const int Pa = (frames[1] + L - 1) / L;
    std::vector<double> noise((size_t)L * Pa * C);
    std::uniform_real_distribution<double> uni(-1.0, 1.0);
    std::vector<T> noise_t(noise.size());
    for (size_t i = 0; i < noise.size(); i++) { noise_t[i] = (T)uni(gen); noise[i] = (double)noise_t[i]; }
    // make the response loud enough to clip (the function reports 0 dB otherwise)
    std::vector<std::vector<T> > loud(C, std::vector<T>(frames[1]));
    std::vector<double> inter((size_t)frames[1] * C);
    for (int f = 0; f < frames[1]; f++) for (int c = 0; c < C; c++) { loud[c][f] = imps[1][c][f] * (T)8; inter[(size_t)f * C + c] = (double)loud[c][f]; }
    ref_vfile_register(L"loud.wav", C, frames[1], rate, inter.data());
    ref_set_noise(noise.data(), (long long)noise.size());
    double att_ref = 0, att = 0;
    int rc = 0, rf = 0, rr = 0;
    if (!ref_calculate_attenuation(L"loud.wav", L, (int)sizeof(T), &att_ref, &rc, &rf, &rr) || rc != C || rf != frames[1] || rr != rate) { printf("%s: reference attenuation failed\n", tag); return 1; }
    if (!preprocessor::calculate_attenuation<T>(loud, L, rate, &att, 0, true, noise_t.data())) { printf("%s: attenuation failed\n", tag); return 1; }
    printf("%s: calculate_attenuation %.9f dB, reference %.9f dB\n", tag, att, att_ref);
    if (!(att_ref < -1.0) || !(std::fabs(att - att_ref) < tol_db)) return 1;
    return 0;
}

int main()
{
    if (run<float>("float", 1e-5, 1e-4)) return 1;
    if (run<double>("double", 1e-12, 1e-9)) return 1;
    printf("preprocessor parity ok\n");
    return 0;
}
