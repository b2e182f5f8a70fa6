// TEST INFRASTRUCTURE ONLY. MSVC/Win32 compatibility layer, force-included (-include) when the
// UNMODIFIED reference sources under /root/reference/brutefir are compiled with g++ into
// oracle/_ref/. It only maps Microsoft CRT names to their POSIX equivalents; it contains no DSP code.
#pragma once
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <errno.h>
#include <wchar.h>
#include <alloca.h>
#include <malloc.h>
#include <stdint.h>
#ifdef __cplusplus
#include <cmath>
#include <string>
#endif

typedef int errno_t;

static inline void *bfir_compat_aligned_malloc(size_t size, size_t alignment)
{
    void *p = NULL;
    if (alignment < sizeof(void *)) alignment = sizeof(void *);
    if (posix_memalign(&p, alignment, size ? size : alignment) != 0) return NULL;
    return p;
}
#define _aligned_malloc(size, alignment) bfir_compat_aligned_malloc((size), (alignment))
#define _aligned_free(p) free(p)
#define _alloca(n) alloca(n)
#ifdef __cplusplus
#define _finite(x) std::isfinite((double)(x))
#else
#define _finite(x) isfinite((double)(x))
#endif

// No wisdom / coefficient files exist in the oracle build: every wide-char open reports ENOENT,
// which the reference treats as "no wisdom yet" (brutefir/fftw_convolver.cpp:86-94).
static inline errno_t _wfopen_s(FILE **stream, const wchar_t *, const wchar_t *)
{
    *stream = NULL;
    errno = ENOENT;
    return ENOENT;
}
static inline errno_t fopen_s(FILE **stream, const char *name, const char *mode)
{
    *stream = fopen(name, mode);
    return *stream ? 0 : errno;
}
#define _fileno fileno

// _aligned_realloc is only reached by the text/raw coefficient FILE loaders (brutefir/coeff.cpp:103-190),
// which the oracle never calls (no files); a size-unaware realloc is enough to link.
static inline void *bfir_compat_aligned_realloc(void *p, size_t size, size_t alignment)
{
    void *q = bfir_compat_aligned_malloc(size, alignment);
    if (p != NULL && q != NULL) {
        size_t old = malloc_usable_size(p);
        memcpy(q, p, old < size ? old : size);
        free(p);
    }
    return q;
}
#define _aligned_realloc(p, size, alignment) bfir_compat_aligned_realloc((p), (size), (alignment))
