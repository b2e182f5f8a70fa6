// Drives the SDK-free plug-in adapter (host/dsp_bfir.hpp) and the offline tools (host/preprocessor.hpp)
// on the GPU engine. "gpu" argument: full run; otherwise only the no-device contract is checked.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include "../../foo-dsp-bfir_b200/host/dsp_bfir.hpp"
#include "../../foo-dsp-bfir_b200/host/preprocessor.hpp"

struct Ctx {
    std::vector<std::vector<double> > h;
    std::vector<void *> ptrs;
    std::vector<float> out;
    size_t chunks;
    bool have_filter;
};

static bool provider(void *user, unsigned channels, unsigned, void ***coeffs, int *n_coeffs, int *length, double *scale)
{
    Ctx *c = (Ctx *)user;
    if (!c->have_filter) return false;
    c->h.assign(channels, std::vector<double>(3000, 0.0));       // 3000 taps -> 3 blocks of 1024
    for (unsigned k = 0; k < channels; k++) c->h[k][100 * (k + 1)] = 0.5;   // delay 100(k+1), gain 0.5
    c->ptrs.resize(channels);
    for (unsigned k = 0; k < channels; k++) c->ptrs[k] = c->h[k].data();
    *coeffs = c->ptrs.data(); *n_coeffs = (int)channels; *length = 3000; *scale = 1.0;
    return true;
}

static void sink(void *user, const audio_sample *data, size_t n, unsigned channels, unsigned)
{
    Ctx *c = (Ctx *)user;
    c->out.insert(c->out.end(), data, data + n * channels);
    c->chunks++;
}

int main(int argc, char **argv)
{
    const bool gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    Ctx ctx; ctx.chunks = 0; ctx.have_filter = false;
    {   // no filter configured: audio passes through (on_chunk returns true), nothing is emitted
        dsp_bfir dsp(provider, sink, &ctx);
        std::vector<float> x(700 * 2, 0.25f);
        if (!dsp.on_chunk(x.data(), 700, 2, 44100) || ctx.chunks != 0) { printf("pass-through contract broken\n"); return 1; }
        if (dsp.get_latency() != 0 || dsp.need_track_change_mark()) return 1;
    }
    if (!gpu) { printf("adapter contract ok (no device)\n"); return 0; }

    ctx.have_filter = true;
    dsp_bfir dsp(provider, sink, &ctx, true);
    const unsigned C = 2;
    std::vector<float> x((size_t)5000 * C);
    for (size_t n = 0; n < 5000; n++) for (unsigned k = 0; k < C; k++) x[n * C + k] = (float)std::sin(0.001 * n * (k + 1));
    // odd chunk sizes: 700 + 1500 + 2800 = 5000 frames -> 4 full blocks emitted, 904 frames stay buffered
    size_t pos = 0, sizes[3] = { 700, 1500, 2800 };
    for (int i = 0; i < 3; i++) { if (dsp.on_chunk(&x[pos * C], sizes[i], C, 44100)) { printf("unexpected pass-through\n"); return 1; } pos += sizes[i]; }
    if (ctx.chunks != 4 || ctx.out.size() != (size_t)4096 * C) { printf("framing wrong: %zu chunks\n", ctx.chunks); return 1; }
    double worst = 0;
    for (size_t n = 0; n < 4096; n++) for (unsigned k = 0; k < C; k++) {
        const size_t d = 100 * (k + 1);
        const double want = n >= d ? 0.5 * x[(n - d) * C + k] : 0.0;
        worst = std::fmax(worst, std::fabs(ctx.out[n * C + k] - want));
    }
    if (worst > 1e-6) { printf("adapter output wrong: %g\n", worst); return 1; }
    dsp.flush();
    // format change -> re-init (new engine), the 904 buffered frames are dropped like the reference
    ctx.out.clear(); ctx.chunks = 0;
    std::vector<float> y((size_t)1024 * 3, 0.1f);
    if (dsp.on_chunk(y.data(), 1024, 3, 48000) || ctx.chunks != 1 || ctx.out.size() != (size_t)1024 * 3) { printf("re-init wrong\n"); return 1; }

    // offline tools: cascade of two delays = one delay (first block only, like the reference)
    std::vector<std::vector<std::vector<double> > > imps(2, std::vector<std::vector<double> >(1, std::vector<double>(256, 0.0)));
    imps[0][0][3] = 1.0; imps[1][0][5] = 2.0;
    std::vector<std::vector<double> > casc;
    if (!preprocessor::convolve_impulses<double>(imps, std::vector<double>(2, 1.0), 256, 2, 44100, casc)) { printf("cascade failed\n"); return 1; }
    for (size_t i = 0; i < casc[0].size(); i++) {
        const double want = i == 8 ? 2.0 : 0.0;
        if (std::fabs(casc[0][i] - want) > 1e-12) { printf("cascade wrong at %zu: %g\n", i, casc[0][i]); return 1; }
    }
    // attenuation probe: gain 4 impulse -> peak just under 4 -> about -12 dB
    std::vector<std::vector<float> > resp(2, std::vector<float>(1024, 0.f));
    resp[0][0] = 4.0f; resp[1][7] = 0.5f;
    double att = 0;
    if (!preprocessor::calculate_attenuation<float>(resp, 512, 44100, &att)) { printf("attenuation failed\n"); return 1; }
    if (!(att < -11.9 && att > -12.05)) { printf("attenuation wrong: %g\n", att); return 1; }
    printf("adapter + tools ok: worst %.3g, attenuation %.3f dB\n", worst, att);
    return 0;
}
