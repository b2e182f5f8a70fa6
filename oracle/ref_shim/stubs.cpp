// TEST INFRASTRUCTURE ONLY. Link-time stand-ins for the parts of the reference that are out of scope
// for the hot path (file I/O through libsndfile, profile-directory paths, the foobar console) so that
// the UNMODIFIED hot-path sources link into oracle/_ref/libbfir_ref.so. No DSP arithmetic lives here.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <utility>
#include <wchar.h>
#include <stdint.h>
#include "pinfo.h"
#include "bfir_path.hpp"
#include "buffer.hpp"

// ---- pinfo (reference brutefir/pinfo.c:26-39 sends it to the foobar console) ----
static void (*g_print_cb)(const char *) = NULL;
extern "C" void set_print_callback(void (*cb)(const char *)) { g_print_cb = cb; }
extern "C" void pinfo(const char *format, ...)
{
    char msg[1024];
    va_list ap;
    va_start(ap, format);
    vsnprintf(msg, sizeof(msg), format, ap);
    va_end(ap);
    if (g_print_cb != NULL) g_print_cb(msg);
    else if (getenv("BFIR_REF_VERBOSE") != NULL) fprintf(stderr, "[ref] %s\n", msg);
}

// ---- profile paths: only used to name the wisdom / cache files, which never exist here ----
namespace bfir_path {
std::wstring append_path(const std::wstring filename) { return L"/nonexistent/bfir/" + filename; }
std::wstring append_temp_path(const std::wstring filename) { return L"/nonexistent/bfir/tmp/" + filename; }
}

// ---- sound-file layer ----
// The equalizer hands its rendered filter to save_to_snd_file (brutefir/equalizer.cpp:284-289); the
// stub keeps the last buffer so the wrapper in ref_capi.cpp can return it instead of a WAV file.
std::vector<unsigned char> g_last_saved;
int g_last_saved_channels = 0, g_last_saved_frames = 0, g_last_saved_realsize = 0;

// Virtual sound files: the offline drivers of brutefir/preprocessor.cpp read their impulse responses through
// buffer::get_snd_file_params / load_from_snd_file (libsndfile in the reference, brutefir/buffer.cpp:38-178). The
// stub serves them from a registry of in-memory "files" the test fills (ref_vfile_register in ref_capi.cpp), with
// the reference's semantics: n_frames clipped to max_frames, zero padding up to max_frames when pad is set, samples
// converted to the engine's precision the way sf_readf_float / sf_readf_double would.
struct vfile { int channels, frames, rate; std::vector<double> data; };   // interleaved
static std::vector<std::pair<std::wstring, vfile> > g_vfiles;
static std::vector<double> g_noise;      // what load_white_noise hands out (the reference seeds from time())
static size_t g_noise_pos = 0;
void stub_vfile_register(const wchar_t *name, int channels, int frames, int rate, const double *data)
{
    vfile f;
    f.channels = channels; f.frames = frames; f.rate = rate;
    f.data.assign(data, data + (size_t)channels * frames);
    for (size_t i = 0; i < g_vfiles.size(); i++) if (g_vfiles[i].first == name) { g_vfiles[i].second = f; return; }
    g_vfiles.push_back(std::make_pair(std::wstring(name), f));
}
void stub_vfile_clear() { g_vfiles.clear(); }
void stub_set_noise(const double *data, size_t n) { g_noise.assign(data, data + n); g_noise_pos = 0; }
static const vfile *find_vfile(const wchar_t *name)
{
    for (size_t i = 0; i < g_vfiles.size(); i++) if (g_vfiles[i].first == name) return &g_vfiles[i].second;
    return NULL;
}

namespace util {
// brutefir/util.cpp:46-56 (the file's string helpers need a locale; only this arithmetic helper is linked)
uint32_t get_next_multiple(uint32_t value, uint32_t factor)
{
    uint32_t multiple = factor;
    while (value > multiple) multiple += factor;
    return multiple;
}
}

namespace buffer {
bool check_snd_file(const wchar_t *, int, int) { return false; }
bool get_snd_file_params(const wchar_t *filename, int *n_channels, int *n_frames, int *sampling_rate)
{
    const vfile *f = find_vfile(filename);
    if (f == NULL) return false;
    *n_channels = f->channels; *n_frames = f->frames; *sampling_rate = f->rate;
    return true;
}
void *load_from_snd_file(const wchar_t *filename, int *n_channels, int *n_frames, int realsize, int max_frames, bool pad)
{
    const vfile *f = find_vfile(filename);
    if (f == NULL) return NULL;
    *n_channels = f->channels;
    *n_frames = max_frames == -1 ? f->frames : (max_frames > f->frames ? f->frames : max_frames);     // buffer.cpp:55
    const size_t alloc = (size_t)(pad ? max_frames : *n_frames) * f->channels;
    void *buf = bfir_compat_aligned_malloc(alloc * realsize, 32);
    memset(buf, 0, alloc * realsize);
    const size_t n = (size_t)*n_frames * f->channels;
    if (realsize == 4) for (size_t i = 0; i < n; i++) ((float *)buf)[i] = (float)f->data[i];
    else for (size_t i = 0; i < n; i++) ((double *)buf)[i] = f->data[i];
    return buf;
}
// buffer.cpp:338-380 de-interleaves through raw2real with the native float format, i.e. a strided copy
void **deinterlace(void *buffer, int n_channels, int n_frames, int realsize)
{
    void **bufs = (void **)bfir_compat_aligned_malloc(n_channels * sizeof(void *), 32);
    for (int c = 0; c < n_channels; c++) {
        bufs[c] = bfir_compat_aligned_malloc((size_t)n_frames * realsize, 32);
        for (int f = 0; f < n_frames; f++)
            memcpy((unsigned char *)bufs[c] + (size_t)f * realsize, (unsigned char *)buffer + ((size_t)f * n_channels + c) * realsize, realsize);
    }
    return bufs;
}
// buffer.cpp:455-493 draws uniform noise in [-1, 1) from a generator seeded with time(); the stub hands out the
// sequence the test registered (ref_set_noise), so both sides of a comparison see the same samples
void *load_white_noise(int n_channels, int n_frames, int realsize)
{
    const size_t n = (size_t)n_channels * n_frames;
    void *buf = bfir_compat_aligned_malloc(n * realsize, 32);
    for (size_t i = 0; i < n; i++) {
        const double v = g_noise.empty() ? 0.0 : g_noise[(g_noise_pos + i) % g_noise.size()];
        if (realsize == 4) ((float *)buf)[i] = (float)v; else ((double *)buf)[i] = v;
    }
    g_noise_pos += n;
    return buf;
}
void *interlace(void **buffers, int n_channels, int n_frames, int realsize)
{
    unsigned char *out = (unsigned char *)malloc((size_t)n_channels * n_frames * realsize);
    for (int f = 0; f < n_frames; f++)
        for (int c = 0; c < n_channels; c++)
            memcpy(out + ((size_t)f * n_channels + c) * realsize,
                   (unsigned char *)buffers[c] + (size_t)f * realsize, realsize);
    return out;
}
void save_to_snd_file(const wchar_t *filename, void *buf, int n_channels, int n_frames, int realsize, int)
{
    g_last_saved.assign((unsigned char *)buf, (unsigned char *)buf + (size_t)n_channels * n_frames * realsize);
    g_last_saved_channels = n_channels;
    g_last_saved_frames = n_frames;
    g_last_saved_realsize = realsize;
    // the equalizer leaks the buffer it hands over (equalizer.cpp:282-289, files named "eq-..."): the stub owns it
    // instead; preprocessor::convolve_impulses frees its own (preprocessor.cpp:206-210, files named "file-...")
    if (wcsstr(filename, L"/eq-") != NULL) free(buf);
}
}
