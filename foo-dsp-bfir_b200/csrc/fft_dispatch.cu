// Dispatch table over the per-size FFT translation units (fft_inst.cu).
#include "fft_dispatch.hpp"

namespace bfir {

#define BFIR_DECL(tag, m)                                                                                      \
    cudaError_t launch_fwd_##tag##_m##m(dim3, cudaStream_t, const FwdArgs &, const void *, int, int);          \
    cudaError_t launch_inv_##tag##_m##m(dim3, cudaStream_t, const InvArgs &, const void *, int, int);
#define BFIR_FOR_F32(X) X(f32, 4) X(f32, 5) X(f32, 6) X(f32, 7) X(f32, 8) X(f32, 9) X(f32, 10) X(f32, 11) X(f32, 12) X(f32, 13) X(f32, 14)
#define BFIR_FOR_F64(X) X(f64, 4) X(f64, 5) X(f64, 6) X(f64, 7) X(f64, 8) X(f64, 9) X(f64, 10) X(f64, 11) X(f64, 12) X(f64, 13)
BFIR_FOR_F32(BFIR_DECL)
BFIR_FOR_F64(BFIR_DECL)

static const int kMinLog2M = 4, kMaxLog2M_f32 = 14, kMaxLog2M_f64 = 13;

#define BFIR_FWD_ENTRY(tag, m) launch_fwd_##tag##_m##m,
#define BFIR_INV_ENTRY(tag, m) launch_inv_##tag##_m##m,
static const fwd_launcher_t kFwdF32[] = { BFIR_FOR_F32(BFIR_FWD_ENTRY) };
static const inv_launcher_t kInvF32[] = { BFIR_FOR_F32(BFIR_INV_ENTRY) };
static const fwd_launcher_t kFwdF64[] = { BFIR_FOR_F64(BFIR_FWD_ENTRY) };
static const inv_launcher_t kInvF64[] = { BFIR_FOR_F64(BFIR_INV_ENTRY) };

bool rfft_supported(int realsize, int log2m)
{
    if (realsize == 4) return log2m >= kMinLog2M && log2m <= kMaxLog2M_f32;
    if (realsize == 8) return log2m >= kMinLog2M && log2m <= kMaxLog2M_f64;
    return false;
}

size_t rfft_smem_bytes(int realsize, int log2m)
{
    const size_t m = (size_t)1 << log2m;
    return (m + (m >> 4)) * 2 * (size_t)realsize;
}

cudaError_t launch_rfft_forward(int realsize, int log2m, dim3 grid, cudaStream_t stream, const FwdArgs &a,
                                const void *tw, int sm, int sn)
{
    if (!rfft_supported(realsize, log2m)) return cudaErrorInvalidValue;
    return (realsize == 4 ? kFwdF32 : kFwdF64)[log2m - kMinLog2M](grid, stream, a, tw, sm, sn);
}

cudaError_t launch_rfft_inverse(int realsize, int log2m, dim3 grid, cudaStream_t stream, const InvArgs &a,
                                const void *tw, int sm, int sn)
{
    if (!rfft_supported(realsize, log2m)) return cudaErrorInvalidValue;
    return (realsize == 4 ? kInvF32 : kInvF64)[log2m - kMinLog2M](grid, stream, a, tw, sm, sn);
}

} // namespace bfir
