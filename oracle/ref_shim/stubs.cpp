// TEST INFRASTRUCTURE ONLY. Link-time stand-ins for the parts of the reference that are out of scope
// for the hot path (file I/O through libsndfile, profile-directory paths, the foobar console) so that
// the UNMODIFIED hot-path sources link into oracle/_ref/libbfir_ref.so. No DSP arithmetic lives here.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
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

namespace buffer {
bool check_snd_file(const wchar_t *, int, int) { return false; }
void *load_from_snd_file(const wchar_t *, int *, int *, int, int, bool) { return NULL; }
void **deinterlace(void *, int, int, int) { return NULL; }
void *interlace(void **buffers, int n_channels, int n_frames, int realsize)
{
    unsigned char *out = (unsigned char *)malloc((size_t)n_channels * n_frames * realsize);
    for (int f = 0; f < n_frames; f++)
        for (int c = 0; c < n_channels; c++)
            memcpy(out + ((size_t)f * n_channels + c) * realsize,
                   (unsigned char *)buffers[c] + (size_t)f * realsize, realsize);
    return out;
}
void save_to_snd_file(const wchar_t *, void *buf, int n_channels, int n_frames, int realsize, int)
{
    g_last_saved.assign((unsigned char *)buf, (unsigned char *)buf + (size_t)n_channels * n_frames * realsize);
    g_last_saved_channels = n_channels;
    g_last_saved_frames = n_frames;
    g_last_saved_realsize = realsize;
    free(buf); // the reference leaks this buffer (equalizer.cpp:282-289); the stub owns it instead
}
}
