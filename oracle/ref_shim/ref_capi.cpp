// TEST INFRASTRUCTURE ONLY. extern "C" handles onto the UNMODIFIED reference classes (fftw_convolver,
// brutefir, dither, equalizer, coeff) compiled from /root/reference/brutefir into
// oracle/_ref/libbfir_ref.so, so that tests/ and bench.py's cpu_baseline / --impl reference can drive
// them through ctypes. This file contains glue only; the only reference logic restated here is the
// private sample-format table (brutefir/brutefir.cpp:436-539), needed to build a buffer_format_t.
#define private public // read-only inspection of reference state (dither table, overflow counters)
#include "global.h"
#include "dither.hpp"
#include "fftw_convolver.hpp"
#include "brutefir.hpp"
#include "coeff.hpp"
#include "equalizer.hpp"
#include "raw2real.hpp"
#include "real2raw.hpp"
#include "preprocessor.hpp"
#undef private
#include <string.h>
#include <stdlib.h>
#include <vector>

void stub_vfile_register(const wchar_t *name, int channels, int frames, int rate, const double *data);
void stub_vfile_clear();
void stub_set_noise(const double *data, size_t n);
extern std::vector<unsigned char> g_last_saved;
extern int g_last_saved_channels, g_last_saved_frames, g_last_saved_realsize;

namespace {

// mirrors brutefir::setup_sample_format + setup_input/setup_output (brutefir.cpp:436-582)
int fill_bf(struct buffer_format_t *bf, int format, bool normalized, int index, int spacing)
{
    struct sample_format_t *sf = &bf->sf;
    memset(bf, 0, sizeof(*bf));
    sf->format = format;
    switch (format) {
    case BF_SAMPLE_FORMAT_S8: sf->bytes = 1; sf->isfloat = false; sf->swap = false; break;
    case BF_SAMPLE_FORMAT_S16_LE: sf->bytes = 2; sf->isfloat = false; sf->swap = false; break;
    case BF_SAMPLE_FORMAT_S16_BE: sf->bytes = 2; sf->isfloat = false; sf->swap = true; break;
    case BF_SAMPLE_FORMAT_S24_LE: sf->bytes = 3; sf->isfloat = false; sf->swap = false; break;
    case BF_SAMPLE_FORMAT_S24_BE: sf->bytes = 3; sf->isfloat = false; sf->swap = true; break;
    case BF_SAMPLE_FORMAT_S32_LE: sf->bytes = 4; sf->isfloat = false; sf->swap = false; break;
    case BF_SAMPLE_FORMAT_S32_BE: sf->bytes = 4; sf->isfloat = false; sf->swap = true; break;
    case BF_SAMPLE_FORMAT_FLOAT_LE: sf->bytes = 4; sf->isfloat = true; sf->swap = false; break;
    case BF_SAMPLE_FORMAT_FLOAT_BE: sf->bytes = 4; sf->isfloat = true; sf->swap = true; break;
    case BF_SAMPLE_FORMAT_FLOAT64_LE: sf->bytes = 8; sf->isfloat = true; sf->swap = false; break;
    case BF_SAMPLE_FORMAT_FLOAT64_BE: sf->bytes = 8; sf->isfloat = true; sf->swap = true; break;
    default: return -1;
    }
    sf->sbytes = sf->bytes;
    if (sf->isfloat) {
        sf->scale = 1.0;
    } else {
        double full = (double)(1 << ((sf->bytes << 3) - 1));
        sf->scale = normalized ? 1.0 / full : full;
    }
    bf->byte_offset = index * sf->bytes;
    bf->sample_spacing = spacing;
    return 0;
}

struct conv_handle {
    fftw_convolver *conv;
    dither *dith;
    std::vector<struct dither_state_t> dstate;
    int n_channels;
};

} // namespace

extern "C" {

// ------------------------------------------------------------------ fftw_convolver
void *ref_conv_new(int length, int realsize, int n_channels, int sample_rate)
{
    if ((realsize != 4 && realsize != 8) || length < 1 || (length & (length - 1)) != 0) return NULL;
    conv_handle *h = new conv_handle;
    h->n_channels = n_channels > 0 ? n_channels : 1;
    h->dstate.resize(h->n_channels);
    h->dith = new dither(h->n_channels, sample_rate, realsize, 0, length, h->dstate.data());
    h->conv = new fftw_convolver(length, realsize, h->dith);
    return h;
}

void ref_conv_delete(void *p)
{
    conv_handle *h = (conv_handle *)p;
    delete h->conv;
    delete h->dith;
    delete h;
}

int ref_conv_cbufsize(void *p) { return ((conv_handle *)p)->conv->convolver_cbufsize(); }

int ref_conv_raw2cbuf(void *p, void *rawbuf, void *cbuf, void *next_cbuf, int format, int index, int spacing)
{
    struct buffer_format_t bf;
    if (fill_bf(&bf, format, true, index, spacing) != 0) return -1;
    ((conv_handle *)p)->conv->convolver_raw2cbuf(rawbuf, cbuf, next_cbuf, &bf, NULL, NULL);
    return 0;
}

void ref_conv_time2freq(void *p, void *in, void *out) { ((conv_handle *)p)->conv->convolver_time2freq(in, out); }
void ref_conv_freq2time(void *p, void *in, void *out) { ((conv_handle *)p)->conv->convolver_freq2time(in, out); }

void ref_conv_mixnscale(void *p, void **in, void *out, double *scales, int n_bufs, int mixmode)
{
    ((conv_handle *)p)->conv->convolver_mixnscale(in, out, scales, n_bufs, mixmode);
}

void ref_conv_convolve_inplace(void *p, void *cbuf, void *coeffs) { ((conv_handle *)p)->conv->convolver_convolve_inplace(cbuf, coeffs); }
void ref_conv_convolve(void *p, void *in, void *coeffs, void *out) { ((conv_handle *)p)->conv->convolver_convolve(in, coeffs, out); }
void ref_conv_convolve_add(void *p, void *in, void *coeffs, void *out) { ((conv_handle *)p)->conv->convolver_convolve_add(in, coeffs, out); }
void ref_conv_crossfade_inplace(void *p, void *in, void *xfade, void *buffer) { ((conv_handle *)p)->conv->convolver_crossfade_inplace(in, xfade, buffer); }
void ref_conv_dirac_convolve(void *p, void *in, void *out) { ((conv_handle *)p)->conv->convolver_dirac_convolve(in, out); }
void ref_conv_dirac_convolve_inplace(void *p, void *cbuf) { ((conv_handle *)p)->conv->convolver_dirac_convolve_inplace(cbuf); }
void ref_conv_convolve_eval(void *p, void *in, void *buffer, void *out) { ((conv_handle *)p)->conv->convolver_convolve_eval(in, buffer, out); }

// small one-shot convolver (fftw_convolver.cpp:698-777). n_coeffs = 1 makes the reference shift by -1
// (log2_roof(1) == -1, log2.h:38-43), so the wrappers refuse it instead of calling into that.
int ref_conv_td_block_length(void *p, int n_coeffs)
{
    if (n_coeffs == 1) return -1;
    return ((conv_handle *)p)->conv->convolver_td_block_length(n_coeffs);
}
void *ref_conv_td_new(void *p, void *coeffs, int n_coeffs)
{
    if (n_coeffs < 2) return NULL;
    return ((conv_handle *)p)->conv->convolver_td_new(coeffs, n_coeffs);
}
void ref_conv_td_coeffs(void *tdc, void *dst, int nbytes) { memcpy(dst, ((td_conv_t *)tdc)->coeffs, nbytes); }
void ref_conv_td_convolve(void *p, void *tdc, void *overlap_block)
{
    ((conv_handle *)p)->conv->convolver_td_convolve((td_conv_t *)tdc, overlap_block);
}
void ref_conv_td_free(void *tdc) // the reference never frees a td_conv_t
{
    if (tdc == NULL) return;
    _aligned_free(((td_conv_t *)tdc)->coeffs);
    free(tdc);
}

// overflow: {n_overflows, intlargest, largest, max} passed as a caller-owned bfoverflow_t image
int ref_conv_cbuf2raw(void *p, void *cbuf, void *outbuf, int format, int index, int spacing,
                      int apply_dither, int dither_channel, struct bfoverflow_t *overflow)
{
    conv_handle *h = (conv_handle *)p;
    struct buffer_format_t bf;
    if (fill_bf(&bf, format, false, index, spacing) != 0) return -1;
    if (dither_channel < 0 || dither_channel >= h->n_channels) return -1;
    h->conv->convolver_cbuf2raw(cbuf, outbuf, &bf, apply_dither != 0, &h->dstate[dither_channel], overflow);
    return 0;
}

// returns 0 on success, -1 when the reference returns NULL (NaN/Inf among coefficients)
int ref_conv_coeffs2cbuf(void *p, void *coeffs, int n_coeffs, double scale, void *dest)
{
    return ((conv_handle *)p)->conv->convolver_coeffs2cbuf(coeffs, n_coeffs, scale, dest) == NULL ? -1 : 0;
}

void ref_conv_runtime_coeffs2cbuf(void *p, void *src, void *dest) { ((conv_handle *)p)->conv->convolver_runtime_coeffs2cbuf(src, dest); }

// ------------------------------------------------------------------ dither inspection
int ref_conv_dither_table_size(void *p) { return ((conv_handle *)p)->dith->dither_randtab_size; }
const int8_t *ref_conv_dither_table(void *p) { return ((conv_handle *)p)->dith->dither_randtab; }
int ref_conv_dither_ptr(void *p, int ch) { return ((conv_handle *)p)->dstate[ch].randtab_ptr; }
// copies randmap[-256 .. 254] (511 entries, realsize bytes each)
void ref_conv_dither_map(void *p, void *out)
{
    conv_handle *h = (conv_handle *)p;
    memcpy(out, h->dith->dither_randmap_ptr, 511 * (size_t)h->dith->realsize);
}

// ------------------------------------------------------------------ raw codecs, called directly
void ref_raw2real(int realsize, void *realbuf, void *rawbuf, int bytes, int shift, int isfloat, int spacing, int swap, int n)
{
    if (realsize == 4) raw2real::raw2realf(realbuf, rawbuf, bytes, shift, isfloat != 0, spacing, swap != 0, n);
    else raw2real::raw2reald(realbuf, rawbuf, bytes, shift, isfloat != 0, spacing, swap != 0, n);
}

// ------------------------------------------------------------------ brutefir engine
void *ref_bfir_new(int filter_length, int filter_blocks, int realsize, int channels, int in_format,
                   int out_format, int sampling_rate, int apply_dither)
{
    if (channels < 1 || channels > BF_MAXCHANNELS) return NULL;
    if ((realsize != 4 && realsize != 8) || filter_length < 1 || (filter_length & (filter_length - 1)) != 0) return NULL;
    return new brutefir(filter_length, filter_blocks, realsize, channels, in_format, out_format,
                        sampling_rate, apply_dither != 0);
}

void ref_bfir_delete(void *p) { delete (brutefir *)p; }
int ref_bfir_is_initialized(void *p) { return ((brutefir *)p)->is_initialized() ? 1 : 0; }

int ref_bfir_set_coeff(void *p, void **coeffs, int n_coeffs, int length, int coeff_blocks, double scale)
{
    return ((brutefir *)p)->set_coeff(coeffs, n_coeffs, length, coeff_blocks, scale);
}

int ref_bfir_run(void *p, void *inbuf, void *outbuf) { return ((brutefir *)p)->run(inbuf, outbuf); }
void ref_bfir_reset(void *p) { ((brutefir *)p)->reset(); }

void ref_bfir_get_overflow(void *p, int ch, struct bfoverflow_t *out) { *out = ((brutefir *)p)->overflow[ch]; }
int ref_bfir_dither_ptr(void *p, int ch) { return ((brutefir *)p)->bfconf->dither_state[ch].randtab_ptr; }
unsigned int ref_bfir_blockcounter(void *p) { return ((brutefir *)p)->blockcounter; }

// ------------------------------------------------------------------ coeff helpers
// runs coeff::preprocess_coeff (coeff.cpp:293-354) and copies the P spectra into dest[P][cbufsize]
int ref_preprocess_coeff(void *p, void *coeffs, int filter_length, int coeff_blocks, int coeff_length,
                         int realsize, double scale, void *dest)
{
    conv_handle *h = (conv_handle *)p;
    void **c = coeff::preprocess_coeff(h->conv, coeffs, filter_length, coeff_blocks, coeff_length, realsize, scale);
    if (c == NULL) return -1;
    int sz = h->conv->convolver_cbufsize(), rc = 0;
    for (int n = 0; n < coeff_blocks; n++) {
        if (c[n] == NULL) { rc = -1; continue; }
        memcpy((unsigned char *)dest + (size_t)n * sz, c[n], sz);
        free(c[n]);
    }
    free(c);
    return rc;
}

// ------------------------------------------------------------------ equalizer
// Runs equalizer::generate (equalizer.cpp:87-140) -> render_f/d -> (stubbed) save_to_snd_file and
// returns channel 0 of the rendered taps/2-sample filter in out[] (realsize bytes per sample).
int ref_equalizer_render(int block_length, int n_blocks, int realsize, int n_channels, int sampling_rate,
                         int n_bands, double *freq, double *mag, double *phase, void *out, int out_len)
{
    int total = block_length * n_blocks;
    if (total < 2 || (total & (total - 1)) != 0 || n_bands > BAND_COUNT) return -1;
    equalizer eq(block_length, n_blocks, realsize, n_channels, sampling_rate);
    g_last_saved.clear();
    eq.generate(n_bands, freq, mag, phase);
    if (g_last_saved.empty()) return -1;
    int frames = g_last_saved_frames < out_len ? g_last_saved_frames : out_len;
    for (int f = 0; f < frames; f++)
        memcpy((unsigned char *)out + (size_t)f * realsize,
               g_last_saved.data() + (size_t)f * g_last_saved_channels * realsize, realsize);
    return frames;
}

extern "C" const char *oracle_fft_provider_name(void);   // fftw_api.cpp / fftw_api_mkl.cpp
const char *ref_fft_provider(void) { return oracle_fft_provider_name(); }

}


// ------------------------------------------------------------------ preprocessor (offline drivers of the block loop)
// The UNMODIFIED brutefir/preprocessor.cpp, fed from in-memory "sound files" (stubs.cpp) instead of libsndfile.
extern "C" {

void ref_vfile_register(const wchar_t *name, int channels, int frames, int rate, const double *interleaved)
{
    stub_vfile_register(name, channels, frames, rate, interleaved);
}
void ref_vfile_clear(void) { stub_vfile_clear(); }
void ref_set_noise(const double *samples, long long n) { stub_set_noise(samples, (size_t)n); }

// preprocessor::convolve_impulses (preprocessor.cpp:34-233). The result "file" is what the function hands to
// save_to_snd_file: g_frames frames x channels, interleaved, in the engine's precision. Returns the number of
// frames (0 on error) and copies at most `capacity` bytes to `out`.
int ref_convolve_impulses(const wchar_t *const *names, const double *scales, int n, int filter_length, int realsize,
                          void *out, long long capacity, int *channels)
{
    std::vector<struct impulse_info> info(n);
    for (int i = 0; i < n; i++) { info[i].filename = names[i]; info[i].scale = scales[i]; }
    g_last_saved.clear();
    g_last_saved_frames = 0;
    std::wstring name = preprocessor::convolve_impulses(info, filter_length, realsize);
    if (name.empty() || g_last_saved.empty()) return 0;
    if (channels) *channels = g_last_saved_channels;
    memcpy(out, g_last_saved.data(), (size_t)capacity < g_last_saved.size() ? (size_t)capacity : g_last_saved.size());
    return g_last_saved_frames;
}

// preprocessor::calculate_attenuation (preprocessor.cpp:250-412); noise from ref_set_noise
int ref_calculate_attenuation(const wchar_t *name, int filter_length, int realsize, double *attenuation, int *n_channels,
                              int *n_frames, int *sampling_rate)
{
    return preprocessor::calculate_attenuation(name, filter_length, realsize, attenuation, n_channels, n_frames, sampling_rate) ? 1 : 0;
}

}
