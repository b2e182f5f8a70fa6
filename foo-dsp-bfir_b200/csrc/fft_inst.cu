// One (precision, size) instantiation of the real-FFT kernels; see fft_dispatch.hpp.
// Compile with -DBFIR_FFT_REAL=float|double -DBFIR_FFT_TAG=f32|f64 -DBFIR_FFT_LOG2M=<4..14>; the variants with
// 8 points per thread add -DBFIR_FFT_LOG2E=3 and their own tag (f64e8).
#include "fft_dispatch.hpp"
#include "eq_kernels.cuh"
#include <cstdlib>

#define BFIR_CAT_(a, b, c, d) a##b##_##c##d
#define BFIR_CAT(a, b, c, d) BFIR_CAT_(a, b, c, d)

namespace bfir {

#ifndef BFIR_FFT_LOG2E
#define BFIR_FFT_LOG2E 4
#endif
typedef BFIR_FFT_REAL real_t;
static constexpr int kLog2E = BFIR_FFT_LOG2E;
static constexpr int kLog2M = BFIR_FFT_LOG2M;
static constexpr int kM = 1 << kLog2M;
static constexpr size_t kSmem = (size_t)fft_smem_elems<kM>::value * sizeof(cpx<real_t>);
// the cluster variants of the one-CTA kernels exist from 256 points per CTA (fft_dispatch.cu: cluster_channels)
static constexpr bool kHasCluster = kLog2M >= 8;

// the CTAs of `cx` neighbouring channels as one thread-block cluster along x (de-interleaving through the cluster)
template <class K, class A>
static cudaError_t launch_clustered(K kernel, dim3 grid, int cx, cudaStream_t stream, const A &a, const void *tw, int sm, int sn)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(kM >> kLog2E); cfg.dynamicSmemBytes = kSmem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cx; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, a, (const cpx<real_t> *)tw, sm, sn);
}

template <int R0>
static cudaError_t launch_fwd(dim3 grid, cudaStream_t stream, const FwdArgs &a, const void *tw, int sm, int sn)
{
    static bool configured = false;
    auto kernel = rfft_forward_kernel<real_t, kLog2M, R0, kLog2E>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    grid.z = R0;
    if (R0 == 4) {   // four CTAs per transform: one thread-block cluster along z (DSMEM exchange in the split step)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(kM >> kLog2E); cfg.dynamicSmemBytes = kSmem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 4;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, a, (const cpx<real_t> *)tw, sm, sn);
    }
    if constexpr (R0 == 1 && kHasCluster) {
        if (a.cluster > 1) {
            static bool configured_cl = false;
            auto kcl = rfft_forward_kernel<real_t, kLog2M, 1, kLog2E, true>;
            if (!configured_cl) {
                cudaError_t e = cudaFuncSetAttribute(kcl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
                if (e != cudaSuccess) return e;
                configured_cl = true;
            }
            return launch_clustered(kcl, grid, a.cluster, stream, a, tw, sm, sn);
        }
    }
    kernel<<<grid, kM >> kLog2E, kSmem, stream>>>(a, (const cpx<real_t> *)tw, sm, sn);
    return cudaGetLastError();
}

template <int R0>
static cudaError_t launch_inv(dim3 grid, cudaStream_t stream, const InvArgs &a, const void *tw, int sm, int sn)
{
    static bool configured = false;
    auto kernel = rfft_inverse_kernel<real_t, kLog2M, R0, kLog2E>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    grid.z = R0;
    if constexpr (R0 == 1 && kHasCluster) {
        if (a.cluster > 1) {
            static bool configured_cl = false;
            auto kcl = rfft_inverse_kernel<real_t, kLog2M, 1, kLog2E, true>;
            if (!configured_cl) {
                cudaError_t e = cudaFuncSetAttribute(kcl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
                if (e != cudaSuccess) return e;
                configured_cl = true;
            }
            return launch_clustered(kcl, grid, a.cluster, stream, a, tw, sm, sn);
        }
    }
    kernel<<<grid, kM >> kLog2E, kSmem, stream>>>(a, (const cpx<real_t> *)tw, sm, sn);
    return cudaGetLastError();
}

// kLog2M is the per-CTA sub-transform size; r0 CTAs cooperate on one buffer of 2^kLog2M * r0 points
// four CTAs per transform only where the size needs it: the largest double-precision sub-transform (2^13 points per
// CTA, 16 points per thread) -> 65536-point real transforms
static constexpr bool kHasR04 = sizeof(real_t) == 8 && kLog2M == 13 && kLog2E == 4;

cudaError_t BFIR_CAT(launch_fwd_, BFIR_FFT_TAG, m, BFIR_FFT_LOG2M)(int r0, dim3 grid, cudaStream_t stream, const FwdArgs &a, const void *tw, int sm, int sn)
{
    if constexpr (kHasR04) { if (r0 == 4) return launch_fwd<4>(grid, stream, a, tw, sm, sn); }
    if (r0 == 4) return cudaErrorInvalidValue;
    return r0 == 2 ? launch_fwd<2>(grid, stream, a, tw, sm, sn) : launch_fwd<1>(grid, stream, a, tw, sm, sn);
}

cudaError_t BFIR_CAT(launch_inv_, BFIR_FFT_TAG, m, BFIR_FFT_LOG2M)(int r0, dim3 grid, cudaStream_t stream, const InvArgs &a, const void *tw, int sm, int sn)
{
    if constexpr (kHasR04) { if (r0 == 4) return launch_inv<4>(grid, stream, a, tw, sm, sn); }
    if (r0 == 4) return cudaErrorInvalidValue;
    return r0 == 2 ? launch_inv<2>(grid, stream, a, tw, sm, sn) : launch_inv<1>(grid, stream, a, tw, sm, sn);
}

#if BFIR_FFT_LOG2E == 4
// inverse complex transform of 2^kLog2M points with strided access (four-step building block)
cudaError_t BFIR_CAT(launch_cfft_inv_, BFIR_FFT_TAG, m, BFIR_FFT_LOG2M)(int batch, cudaStream_t stream, const CfftArgs &a)
{
    static bool configured = false;
    auto kernel = cfft_strided_kernel<real_t, kLog2M, true>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kernel<<<batch, kM / 16, kSmem, stream>>>(a);
    return cudaGetLastError();
}
#endif

} // namespace bfir
