// Offline drivers of the block loop, counterparts of the reference's namespace `preprocessor`
// (brutefir/preprocessor.cpp) without its file I/O: the impulse-response cascade of
// convolve_impulses (:34-233) and the white-noise peak probe of calculate_attenuation (:250-412).
// Both only drive class brutefir (host/brutefir.hpp), exactly like the reference.
//
// Reference quirk kept on purpose: both functions hand `filter_length` (one block!) as the coefficient
// length to set_coeff (:170-178, :310), so only the first block of the response is ever convolved;
// `first_block_only = false` uses the whole response instead.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include "brutefir.hpp"

namespace preprocessor
{
    // planar [channel][frames] <-> interleaved helpers (buffer::deinterlace / interlace, float or double)
    template <class T> inline void interleave(const std::vector<std::vector<T> > &planar, std::vector<T> &out)
    {
        const size_t c = planar.size(), n = c ? planar[0].size() : 0;
        out.resize(c * n);
        for (size_t f = 0; f < n; f++) for (size_t k = 0; k < c; k++) out[f * c + k] = planar[k][f];
    }

    // convolve_impulses (:100-233): start from a dirac, stream impulse k through the engine block by
    // block, then make the OUTPUT the coefficient set for impulse k+1. impulses[k][channel][frame],
    // zero-padded to filter_length*filter_blocks frames like load_from_snd_file(..., pad = true).
    template <class T>
    inline bool convolve_impulses(const std::vector<std::vector<std::vector<T> > > &impulses, const std::vector<double> &scales,
                                  int filter_length, int filter_blocks, int sampling_rate, std::vector<std::vector<T> > &result,
                                  bool first_block_only = true)
    {
        const int realsize = (int)sizeof(T);
        const int channels = impulses.empty() ? 0 : (int)impulses[0].size();
        if (channels < 1) return false;
        const int fmt = realsize == 4 ? BFIR_SAMPLE_FORMAT_FLOAT_LE : BFIR_SAMPLE_FORMAT_FLOAT64_LE;
        brutefir filter(filter_length, filter_blocks, realsize, channels, fmt, fmt, sampling_rate, false);   // :103-111
        const size_t frames = (size_t)filter_length * filter_blocks;
        std::vector<T> outbuf(frames * channels, (T)0), inbuf;
        std::vector<std::vector<T> > coeffs(channels, std::vector<T>(filter_length, (T)0));
        for (int c = 0; c < channels; c++) coeffs[c][0] = (T)1;                                               // coeff::load_dirac_coeff, :120
        std::vector<void *> ptrs(channels);
        for (int c = 0; c < channels; c++) ptrs[c] = coeffs[c].data();
        if (filter.set_coeff(ptrs.data(), channels, filter_length, filter_blocks, 1.0) != 0) return false;    // :121
        for (size_t k = 0; k < impulses.size(); k++) {
            if ((int)impulses[k].size() != channels) return false;                                            // :135-139
            std::vector<std::vector<T> > padded(channels, std::vector<T>(frames, (T)0));
            for (int c = 0; c < channels; c++)
                memcpy(padded[c].data(), impulses[k][c].data(), sizeof(T) * std::min(frames, impulses[k][c].size()));
            interleave(padded, inbuf);
            bool status = true;
            for (int n = 0; n < filter_blocks; n++)                                                           // :143-148
                status &= filter.run(&inbuf[(size_t)n * filter_length * channels], &outbuf[(size_t)n * filter_length * channels]) == 0;
            if (!status) return false;
            coeffs.assign(channels, std::vector<T>(frames));                                                  // buffer::deinterlace, :171
            for (size_t f = 0; f < frames; f++) for (int c = 0; c < channels; c++) coeffs[c][f] = outbuf[f * channels + c];
            for (int c = 0; c < channels; c++) ptrs[c] = coeffs[c].data();
            const int length = first_block_only ? filter_length : (int)frames;                                // :176 passes filter_length
            if (filter.set_coeff(ptrs.data(), channels, length, filter_blocks, k < scales.size() ? scales[k] : 1.0) != 0) return false;
        }
        result = coeffs;
        return true;
    }

    // calculate_attenuation (:250-412): uniform white noise in [-1, 1) of filter_length*filter_blocks
    // frames through the response, attenuation = -20 log10(max |y|) when the peak exceeds 1. The
    // reference seeds its generator from time() (buffer.hpp:19); here the seed is a parameter. The peak
    // is read from the engine's overflow statistics (bfoverflow_t.largest, reduced on the device)
    // instead of scanning the output on the host.
    // `noise` (optional): filter_length*filter_blocks*channels interleaved samples to use instead of the seeded
    // generator (what buffer::load_white_noise returned in a run that is being reproduced).
    template <class T>
    inline bool calculate_attenuation(const std::vector<std::vector<T> > &response, int filter_length, int sampling_rate,
                                      double *attenuation, unsigned seed = 0xB200u, bool first_block_only = true, const T *noise = nullptr)
    {
        *attenuation = 0;
        const int realsize = (int)sizeof(T), channels = (int)response.size();
        if (channels < 1) return false;
        const int n_frames = (int)response[0].size();
        const int filter_blocks = (n_frames + filter_length - 1) / filter_length;                             // :279-280
        const int fmt = realsize == 4 ? BFIR_SAMPLE_FORMAT_FLOAT_LE : BFIR_SAMPLE_FORMAT_FLOAT64_LE;
        brutefir filter(filter_length, filter_blocks, realsize, channels, fmt, fmt, sampling_rate, false);
        std::vector<void *> ptrs(channels);
        for (int c = 0; c < channels; c++) ptrs[c] = (void *)response[c].data();
        const int length = first_block_only ? std::min(filter_length, n_frames) : n_frames;                   // :310 passes filter_length
        if (filter.set_coeff(ptrs.data(), channels, length, filter_blocks, 1.0) != 0) return false;
        std::mt19937 gen(seed);
        std::uniform_real_distribution<double> uni(-1.0, 1.0);                                                // buffer.cpp:455-493
        std::vector<T> inbuf((size_t)filter_length * channels), outbuf((size_t)filter_length * channels);
        for (int n = 0; n < filter_blocks; n++) {                                                             // :329-356
            if (noise != nullptr) memcpy(inbuf.data(), noise + (size_t)n * inbuf.size(), sizeof(T) * inbuf.size());
            else for (size_t i = 0; i < inbuf.size(); i++) inbuf[i] = (T)uni(gen);
            if (filter.run(inbuf.data(), outbuf.data()) != 0) return false;
        }
        double peak = 0;
        for (int c = 0; c < channels; c++) {
            bfir_overflow_t ov;
            if (bfir_get_overflow(filter.handle(), c, &ov) != BFIR_OK) return false;
            if (ov.largest > peak) peak = ov.largest;
        }
        if (peak > 1) *attenuation = -20.0 * std::log10(peak);                                                // TO_DB, :361-375
        return true;
    }
}
