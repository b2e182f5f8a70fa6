// C++ host mirror of the reference's `class fftw_convolver` (brutefir/fftw_convolver.hpp:28-166) over
// the C ABI: same method names and argument order. cbufs are DEVICE buffers (convolver_alloc /
// convolver_upload / convolver_download replace the caller-side _aligned_malloc / memcpy of the
// reference, brutefir.cpp:768). buffer_format_t is reduced to the three fields the codecs read.
#pragma once
#include <cstddef>
#include "../../include/bfir_b200.h"

#define CONVOLVER_MIXMODE_INPUT BFIR_MIXMODE_INPUT
#define CONVOLVER_MIXMODE_INPUT_ADD BFIR_MIXMODE_INPUT_ADD
#define CONVOLVER_MIXMODE_OUTPUT BFIR_MIXMODE_OUTPUT

struct bfir_buffer_format_t { int format; int sample_spacing; int byte_offset; }; // global.h:49-54

class fftw_convolver
{
public:
    fftw_convolver(int length, int realsize, int n_dither_channels = 1, int sampling_rate = 44100) : m_c(NULL)
    {
        if (bfir_conv_create(&m_c, length, realsize, n_dither_channels, sampling_rate) != BFIR_OK) m_c = NULL;
    }
    ~fftw_convolver() { bfir_conv_destroy(m_c); }
    bool ok() const { return m_c != NULL; }

    void *convolver_alloc(size_t bytes) { return bfir_conv_alloc(m_c, bytes); }
    void convolver_free(void *p) { bfir_conv_free(m_c, p); }
    int convolver_upload(void *d, const void *h, size_t n) { return bfir_conv_upload(m_c, d, h, n); }
    int convolver_download(void *h, const void *d, size_t n) { return bfir_conv_download(m_c, h, d, n); }

    void convolver_raw2cbuf(void *rawbuf, void *cbuf, void *next_cbuf, struct bfir_buffer_format_t *bf)
    { bfir_conv_raw2cbuf(m_c, rawbuf, cbuf, next_cbuf, bf->format, bf->byte_offset, bf->sample_spacing); }
    void convolver_time2freq(void *input_cbuf, void *output_cbuf) { bfir_conv_time2freq(m_c, input_cbuf, output_cbuf); }
    void convolver_mixnscale(void *input_cbufs[], void *output_cbuf, double scales[], int n_bufs, int mixmode)
    { bfir_conv_mixnscale(m_c, input_cbufs, output_cbuf, scales, n_bufs, mixmode); }
    void convolver_convolve_inplace(void *cbuf, void *coeffs) { bfir_conv_convolve_inplace(m_c, cbuf, coeffs); }
    void convolver_convolve(void *input_cbuf, void *coeffs, void *output_cbuf) { bfir_conv_convolve(m_c, input_cbuf, coeffs, output_cbuf); }
    void convolver_crossfade_inplace(void *input_cbuf, void *crossfade_cbuf, void *buffer_cbuf)
    { bfir_conv_crossfade_inplace(m_c, input_cbuf, crossfade_cbuf, buffer_cbuf); }
    void convolver_convolve_add(void *input_cbuf, void *coeffs, void *output_cbuf) { bfir_conv_convolve_add(m_c, input_cbuf, coeffs, output_cbuf); }
    void convolver_dirac_convolve(void *input_cbuf, void *output_cbuf) { bfir_conv_dirac_convolve(m_c, input_cbuf, output_cbuf); }
    void convolver_dirac_convolve_inplace(void *cbuf) { bfir_conv_dirac_convolve_inplace(m_c, cbuf); }
    void convolver_freq2time(void *input_cbuf, void *output_cbuf) { bfir_conv_freq2time(m_c, input_cbuf, output_cbuf); }
    void convolver_convolve_eval(void *input_cbuf, void *buffer_cbuf, void *output_cbuf)
    { bfir_conv_convolve_eval(m_c, input_cbuf, buffer_cbuf, output_cbuf); }
    // td_conv_t (fftw_convolver.hpp:18-26, :158-166): opaque here; `coeffs` host samples, `overlap_block` device
    int convolver_td_block_length(int n_coeffs) { return bfir_conv_td_block_length(n_coeffs); }
    bfir_td_conv *convolver_td_new(void *coeffs, int n_coeffs)
    {
        bfir_td_conv *t = nullptr;
        return bfir_conv_td_new(m_c, &t, coeffs, n_coeffs) == BFIR_OK ? t : nullptr;
    }
    void convolver_td_convolve(bfir_td_conv *tdc, void *overlap_block) { bfir_conv_td_convolve(m_c, tdc, overlap_block); }
    void convolver_cbuf2raw(void *cbuf, void *outbuf, struct bfir_buffer_format_t *bf, bool apply_dither,
                            int dither_channel, bfir_overflow_t *overflow)
    { bfir_conv_cbuf2raw(m_c, cbuf, outbuf, bf->format, bf->byte_offset, bf->sample_spacing, apply_dither ? 1 : 0, dither_channel, overflow); }
    int convolver_cbufsize(void) { return bfir_conv_cbufsize(m_c); }
    // returns optional_dest, or NULL on NaN/Inf among the coefficients like the reference
    void *convolver_coeffs2cbuf(void *coeffs, int n_coeffs, double scale, void *optional_dest)
    {
        void *dest = optional_dest != NULL ? optional_dest : convolver_alloc((size_t)convolver_cbufsize());
        if (bfir_conv_coeffs2cbuf(m_c, coeffs, n_coeffs, scale, dest) != BFIR_OK) {
            if (optional_dest == NULL) convolver_free(dest);
            return NULL;
        }
        return dest;
    }
    void convolver_runtime_coeffs2cbuf(void *src, void *dest) { bfir_conv_runtime_coeffs2cbuf(m_c, src, dest); }
    bfir_conv *handle() { return m_c; }

private:
    fftw_convolver(const fftw_convolver &);
    fftw_convolver &operator=(const fftw_convolver &);
    bfir_conv *m_c;
};
