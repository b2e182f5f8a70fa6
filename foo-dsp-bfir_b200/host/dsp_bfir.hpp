// Portable counterpart of the plug-in's `class dsp_bfir : dsp_impl_base`
// (foo_dsp_bfir/foo_dsp_bfir.cpp:76-412) without the foobar2000 SDK: the chunk framing contract of
// on_chunk -- accumulate FILTER_LEN frames, run the engine once per full block, emit one chunk per
// block, re-initialise on a channel-count / sample-rate change, pass audio through while no filter is
// available -- over the GPU engine (class brutefir of host/brutefir.hpp).
//
// What the SDK provides is replaced by two callbacks:
//   filter_provider(channels, srate, &coeffs, &n_coeffs, &length, &scale) -> true when a filter exists
//       (the reference builds it from the EQ / impulse files, foo_dsp_bfir.cpp:150-262); coeffs are
//       planar double arrays owned by the provider until the call returns
//   chunk_sink(data, sample_count, channels, srate)  == insert_chunk + set_data_32 (:334-335)
#pragma once
#include <cstdlib>
#include <cstring>
#include "brutefir.hpp"

#define BFIR_FILTER_LEN 1024 // FILTER_LEN, foo_dsp_bfir/common.h:17
#define BFIR_REALSIZE 8      // REALSIZE,   foo_dsp_bfir/common.h:19

typedef float audio_sample;  // set_data_32 consumes float (foo_dsp_bfir.cpp:335)

class dsp_bfir
{
public:
    typedef bool (*filter_provider_t)(void *user, unsigned channels, unsigned srate, void ***coeffs, int *n_coeffs, int *length, double *scale);
    typedef void (*chunk_sink_t)(void *user, const audio_sample *data, size_t sample_count, unsigned channels, unsigned srate);

    dsp_bfir(filter_provider_t provider, chunk_sink_t sink, void *user, bool check_overflows = false)
        : m_provider(provider), m_sink(sink), m_user(user), m_check_overflows(check_overflows),
          m_filter(NULL), m_channels(0), m_srate(0), m_buffer_count(0), m_inbuf(NULL), m_outbuf(NULL), m_pinned(false) {}

    ~dsp_bfir() { delete m_filter; free_buffers(); }

    // foo_dsp_bfir.cpp:100-362. Returns true = "pass the original chunk through" (no filter),
    // false = "the chunk was consumed and replaced by the emitted ones".
    bool on_chunk(const audio_sample *data, size_t sample_count, unsigned channels, unsigned srate)
    {
        bool first_init = false, re_init = false;
        if (channels != m_channels) { if (m_channels == 0) first_init = true; else re_init = true; m_channels = channels; } // :112-124
        if (srate != m_srate) { if (m_srate == 0) first_init = true; else re_init = true; m_srate = srate; }                // :126-138
        if (first_init || re_init) {                                                                                       // :140-300
            delete m_filter;
            m_filter = NULL;
            m_buffer_count = 0;
            void **coeffs = NULL;
            int n_coeffs = 0, length = 0;
            double scale = 1.0;
            if (m_provider != NULL && m_provider(m_user, m_channels, m_srate, &coeffs, &n_coeffs, &length, &scale)) {
                const int blocks = (length + BFIR_FILTER_LEN - 1) / BFIR_FILTER_LEN;          // util::get_next_multiple, :271-272
                m_filter = new brutefir(BFIR_FILTER_LEN, blocks > 0 ? blocks : 1, BFIR_REALSIZE, (int)m_channels,
                                        BFIR_SAMPLE_FORMAT_FLOAT_LE, BFIR_SAMPLE_FORMAT_FLOAT_LE, (int)m_srate, false); // :279-286
                m_filter->set_coeff(coeffs, n_coeffs, length, blocks > 0 ? blocks : 1, scale);                        // :289
                alloc_buffers((size_t)BFIR_FILTER_LEN * m_channels);                                                  // :292-294
            }
        }
        if (m_filter == NULL) return true;                       // :352-357 pass-through
        if (!m_filter->is_initialized()) return false;           // :305-306: chunk dropped, nothing emitted
        const audio_sample *src = data;
        while (sample_count) {                                   // :311-349
            size_t todo = BFIR_FILTER_LEN - m_buffer_count;
            if (todo > sample_count) todo = sample_count;
            memcpy(m_inbuf + m_buffer_count * m_channels, src, todo * m_channels * sizeof(audio_sample));
            src += todo * m_channels;
            sample_count -= todo;
            m_buffer_count += todo;
            if (m_buffer_count == BFIR_FILTER_LEN) {
                if (m_filter->run(m_inbuf, m_outbuf) == 0) {
                    if (m_sink != NULL) m_sink(m_user, m_outbuf, m_buffer_count, m_channels, m_srate);
                    if (m_check_overflows) m_filter->check_overflows();
                }                                                // else: "Filter processing error.", block dropped
                m_buffer_count = 0;
            }
        }
        return false;
    }

    void on_endofplayback() {}                                   // :364 (the tail is never flushed)
    void on_endoftrack() {}                                      // :365
    void flush() { m_buffer_count = 0; }                         // :367-370
    double get_latency() { return 0; }                           // :372-375
    bool need_track_change_mark() { return false; }              // :377-380
    brutefir *filter() { return m_filter; }

private:
    dsp_bfir(const dsp_bfir &);
    dsp_bfir &operator=(const dsp_bfir &);
    filter_provider_t m_provider;
    chunk_sink_t m_sink;
    void *m_user;
    bool m_check_overflows;
    brutefir *m_filter;
    unsigned m_channels, m_srate;
    size_t m_buffer_count;
    // the block buffers are page-locked (bfir_host_alloc) so that bfir_run's copies go straight over the link;
    // plain memory is the fall-back when that fails (bfir_run accepts it)
    audio_sample *m_inbuf, *m_outbuf;
    bool m_pinned;
    void free_buffers()
    {
        if (m_pinned) { bfir_host_free(m_inbuf); bfir_host_free(m_outbuf); } else { free(m_inbuf); free(m_outbuf); }
        m_inbuf = m_outbuf = NULL;
    }
    void alloc_buffers(size_t n)
    {
        free_buffers();
        m_inbuf = (audio_sample *)bfir_host_alloc(n * sizeof(audio_sample));
        m_outbuf = (audio_sample *)bfir_host_alloc(n * sizeof(audio_sample));
        m_pinned = m_inbuf != NULL && m_outbuf != NULL;
        if (!m_pinned) {
            bfir_host_free(m_inbuf); bfir_host_free(m_outbuf);
            m_inbuf = (audio_sample *)malloc(n * sizeof(audio_sample));
            m_outbuf = (audio_sample *)malloc(n * sizeof(audio_sample));
        }
        memset(m_inbuf, 0, n * sizeof(audio_sample));
        memset(m_outbuf, 0, n * sizeof(audio_sample));
    }
};
