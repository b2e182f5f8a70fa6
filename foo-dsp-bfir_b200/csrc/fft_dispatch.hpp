// Run-time dispatch onto the (precision, size) instantiations of the real-FFT kernels. Each
// instantiation is compiled in its own translation unit (fft_inst.cu with -DBFIR_FFT_REAL /
// -DBFIR_FFT_LOG2M) so the build parallelises.
#pragma once
#include "rfft_kernels.cuh"

namespace bfir {

// block length L = M = 2^log2m. Supported: 16 <= L <= 16384 (float), 16 <= L <= 8192 (double):
// the CTA-resident transform needs M complex values (+1/16 padding) in shared memory (<= 227 KB).
bool rfft_supported(int realsize, int log2m);
size_t rfft_smem_bytes(int realsize, int log2m);

// grid = (buffers/channels, partitions); block size and shared memory are implied by the size
cudaError_t launch_rfft_forward(int realsize, int log2m, dim3 grid, cudaStream_t stream, const FwdArgs &a,
                                const void *tw, int tw_shift_m, int tw_shift_n);
cudaError_t launch_rfft_inverse(int realsize, int log2m, dim3 grid, cudaStream_t stream, const InvArgs &a,
                                const void *tw, int tw_shift_m, int tw_shift_n);

typedef cudaError_t (*fwd_launcher_t)(dim3, cudaStream_t, const FwdArgs &, const void *, int, int);
typedef cudaError_t (*inv_launcher_t)(dim3, cudaStream_t, const InvArgs &, const void *, int, int);

} // namespace bfir
