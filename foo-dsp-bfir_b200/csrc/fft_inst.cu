// One (precision, size) instantiation of the real-FFT kernels; see fft_dispatch.hpp.
// Compile with -DBFIR_FFT_REAL=float|double -DBFIR_FFT_TAG=f32|f64 -DBFIR_FFT_LOG2M=<4..14>.
#include "fft_dispatch.hpp"

#define BFIR_CAT_(a, b, c, d) a##b##_##c##d
#define BFIR_CAT(a, b, c, d) BFIR_CAT_(a, b, c, d)

namespace bfir {

typedef BFIR_FFT_REAL real_t;
static constexpr int kLog2M = BFIR_FFT_LOG2M;
static constexpr int kM = 1 << kLog2M;
static constexpr size_t kSmem = (size_t)fft_smem_elems<kM>::value * sizeof(cpx<real_t>);

cudaError_t BFIR_CAT(launch_fwd_, BFIR_FFT_TAG, m, BFIR_FFT_LOG2M)(dim3 grid, cudaStream_t stream, const FwdArgs &a, const void *tw, int sm, int sn)
{
    static bool configured = false;
    auto kernel = rfft_forward_kernel<real_t, kLog2M>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kernel<<<grid, kM / 16, kSmem, stream>>>(a, (const cpx<real_t> *)tw, sm, sn);
    return cudaGetLastError();
}

cudaError_t BFIR_CAT(launch_inv_, BFIR_FFT_TAG, m, BFIR_FFT_LOG2M)(dim3 grid, cudaStream_t stream, const InvArgs &a, const void *tw, int sm, int sn)
{
    static bool configured = false;
    auto kernel = rfft_inverse_kernel<real_t, kLog2M>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kernel<<<grid, kM / 16, kSmem, stream>>>(a, (const cpx<real_t> *)tw, sm, sn);
    return cudaGetLastError();
}

} // namespace bfir
