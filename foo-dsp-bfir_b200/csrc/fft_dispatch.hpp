// Run-time dispatch onto the (precision, size) instantiations of the real-FFT kernels. Each
// instantiation is compiled in its own translation unit (fft_inst.cu with -DBFIR_FFT_REAL /
// -DBFIR_FFT_LOG2M) so the build parallelises.
#pragma once
#include "rfft_kernels.cuh"
#include "eq_kernels.cuh"

namespace bfir {

// block length L = M = 2^log2m. One transform is computed by r0 = 1, 2 or 4 CTAs (rfft_kernels.cuh); each
// CTA holds M/r0 complex values (+1/16 padding) in shared memory (<= 227 KB), which bounds the
// per-CTA size at 2^14 (float) / 2^13 (double). Supported: 16 <= L <= 32768 in both precisions (double at L = 32768:
// four CTAs, the forward kernel as a thread-block cluster).
bool rfft_supported(int realsize, int log2m);
// 1, 2 or 4 CTAs per transform: 2 / 4 when the size needs it, 2 also when there are too few buffers to fill the GPU
int rfft_choose_r0(int realsize, int log2m, long long n_buffers);

// grid = (buffers/channels, partitions); block size, grid.z and shared memory are implied by size and r0.
// tw = table exp(-2 pi i j / N), N = 2 * 2^log2m.
cudaError_t launch_rfft_forward(int realsize, int log2m, int r0, dim3 grid, cudaStream_t stream, const FwdArgs &a, const void *tw);
cudaError_t launch_rfft_inverse(int realsize, int log2m, int r0, dim3 grid, cudaStream_t stream, const InvArgs &a, const void *tw);

// batched inverse complex transform of 2^log2m points (4 <= log2m <= 14 float / 13 double), grid = batch
cudaError_t launch_cfft_inverse(int realsize, int log2m, int batch, cudaStream_t stream, const CfftArgs &a);
int cfft_max_log2m(int realsize);
typedef cudaError_t (*cfft_launcher_t)(int, cudaStream_t, const CfftArgs &);

typedef cudaError_t (*fwd_launcher_t)(int, dim3, cudaStream_t, const FwdArgs &, const void *, int, int);
typedef cudaError_t (*inv_launcher_t)(int, dim3, cudaStream_t, const InvArgs &, const void *, int, int);

} // namespace bfir
