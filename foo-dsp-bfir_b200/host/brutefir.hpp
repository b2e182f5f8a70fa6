// C++ host mirror of the reference's `class brutefir` (brutefir/brutefir.hpp:15-52) over the C ABI of
// libbfir_b200.so: same constructor arguments, same method names, same return codes. A plug-in built
// against the reference (foo_dsp_bfir/foo_dsp_bfir.cpp:279-345) switches engines by including this
// header instead of brutefir/brutefir.hpp; see INTEGRATION.md.
#pragma once
#include <cstddef>
#include "../../include/bfir_b200.h"

class brutefir
{
public:
    brutefir(int filter_length, int filter_blocks, int realsize, int channels, int in_format, int out_format,
             int sampling_rate, bool apply_dither)
        : m_engine(NULL)
    {
        // like the reference constructor (brutefir.cpp:21-44) this cannot throw: on failure the object
        // exists but run() refuses and bfir_last_error() tells why
        if (bfir_create(&m_engine, filter_length, filter_blocks, realsize, channels, in_format, out_format,
                        sampling_rate, apply_dither ? 1 : 0) != BFIR_OK)
            m_engine = NULL;
    }

    ~brutefir() { bfir_destroy(m_engine); }

    bool is_initialized() { return m_engine != NULL && bfir_is_initialized(m_engine) != 0; }

    // set_coeff(const wchar_t *filename, ...) (brutefir.cpp:89-165) needs libsndfile and is out of
    // scope: load the file on the host and pass the planar arrays to the overload below.
    int set_coeff(void **coeffs, int n_coeffs, int length, int coeff_blocks, double scale)
    {
        if (m_engine == NULL) return -2;
        const int rc = bfir_set_coeff(m_engine, (const void *const *)coeffs, n_coeffs, length, coeff_blocks, scale);
        return rc == BFIR_OK ? 0 : -2;
    }

    int run(void *inbuf, void *outbuf)
    {
        if (m_engine == NULL) return -1;
        return bfir_run(m_engine, inbuf, outbuf) == BFIR_OK ? 0 : -1;
    }

    // throughput variant for batch callers (no reference counterpart): pinned buffers, see bfir_run_async
    long long run_async(void *inbuf, void *outbuf) { return m_engine == NULL ? -1 : bfir_run_async(m_engine, inbuf, outbuf); }
    int wait(long long ticket) { return (m_engine != NULL && bfir_wait(m_engine, ticket) == BFIR_OK) ? 0 : -1; }

    void reset() { if (m_engine != NULL) bfir_reset(m_engine); }
    void check_overflows() { if (m_engine != NULL) bfir_check_overflows(m_engine); }

    bfir_engine *handle() { return m_engine; }

private:
    brutefir(const brutefir &);
    brutefir &operator=(const brutefir &);
    bfir_engine *m_engine;
};
