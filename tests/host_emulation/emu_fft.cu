// CPU emulation of the CTA-wide real FFT kernels: every phase of rfft_forward_kernel /
// rfft_inverse_kernel is executed for all threads in turn (a __syncthreads boundary = end of a loop
// over t), using the very same __host__ __device__ functions the GPU runs. Checked against the
// oracle FFT (oracle/fft_r2r.hpp). Runs without a GPU; used by tests/test_host_emulation.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../../foo-dsp-bfir_b200/csrc/rfft_kernels.cuh"
#include "../../oracle/fft_r2r.hpp"

using namespace bfir;

template <class T> static std::vector<cpx<T>> make_tw(int N)
{
    std::vector<cpx<T>> tw(N);
    for (int j = 0; j < N; j++) {
        long double a = -2.0L * M_PIl * j / N;
        tw[j].x = (T)cosl(a); tw[j].y = (T)sinl(a);
    }
    return tw;
}

template <class T> static double rel_rms(const std::vector<T> &a, const std::vector<T> &b)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) { double d = (double)a[i] - (double)b[i]; num += d * d; den += (double)b[i] * (double)b[i]; }
    return std::sqrt(num / (den > 0 ? den : 1));
}

// LOG2MS = sub-transform size per CTA, R0 = CTAs per buffer; total M = MS * R0
template <class T, int LOG2MS, int R0> static int check(double tol)
{
    constexpr int MS = 1 << LOG2MS, M = MS * R0, N = 2 * M, NT = MS / 16;
    typedef cpx<T> C;
    std::mt19937 rng(1234 + LOG2MS + 100 * R0);
    std::uniform_real_distribution<double> u(-1, 1);
    std::vector<T> x(N), hc(N), ord(N), ref(N), back(N);
    for (auto &v : x) v = (T)u(rng);
    auto tw = make_tw<T>(N);
    const int sm = R0 == 2 ? 2 : 1; // log2(N / MS)
    std::vector<C> smem(fft_smem_elems<MS>::value);
    std::vector<C> regs((size_t)NT * 16);
    C (*vs)[16] = reinterpret_cast<C (*)[16]>(regs.data());
    int fails = 0;

    for (int layout = 0; layout < 2; layout++) {
        FwdArgs a = {};
        a.in_mode = IN_TIME; a.out_layout = layout; a.in = x.data(); a.out = layout == LAYOUT_HC ? hc.data() : ord.data();
        a.scale_in = 1.0; a.scale_out = layout == LAYOUT_ORD ? 0.5 : 1.0;
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) fwd_load<T, LOG2MS, R0>(t, 0, 0, r, vs[t], tw.data(), 0, a);
            fft_passes<T, LOG2MS, false, 0, 0>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) BlockFFT<T, LOG2MS, false>::store_natural(t, vs[t], smem.data());
            for (int t = 0; t < NT; t++) fwd_split_store<T, LOG2MS, R0>(t, 0, 0, r, smem.data(), tw.data(), 0, a);
        }
    }
    oracle_fft::RealFFT<T> of(N);
    of.r2hc(x.data(), ref.data());
    double e1 = rel_rms(hc, ref);
    // ORD vs HC consistency (scale 0.5)
    std::vector<T> ord_ref(N);
    for (int k = 0; k < M; k++) {
        int base = ((k >> 2) << 3) + (k & 3);
        ord_ref[base] = ref[k] * (T)0.5;
        ord_ref[base + 4] = (k == 0 ? ref[M] : ref[N - k]) * (T)0.5;
    }
    double e2 = rel_rms(ord, ord_ref);

    // inverse from ORD (scale 2 undoes the 0.5) and from HC
    double e3[2];
    for (int layout = 0; layout < 2; layout++) {
        InvArgs b = {};
        b.in_layout = layout; b.out_mode = OUT_TIME; b.in = layout == LAYOUT_HC ? ref.data() : ord_ref.data();
        b.scale_in = layout == LAYOUT_ORD ? 2.0 : 1.0; b.out = back.data();
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) inv_load<T, LOG2MS, R0>(t, 0, r, vs[t], tw.data(), 0, b);
            fft_passes<T, LOG2MS, true, 0, 0>::run_host(vs, smem.data(), tw.data(), sm);
            OverflowAcc acc = {};
            for (int t = 0; t < NT; t++) inv_store<T, LOG2MS, R0>(t, 0, r, vs[t], b, acc);
        }
        std::vector<T> want(N);
        for (int i = 0; i < N; i++) want[i] = x[i] * (T)N;
        e3[layout] = rel_rms(back, want);
    }
    bool ok = e1 < tol && e2 < tol && e3[0] < tol && e3[1] < tol;
    printf("%s log2ms=%2d R0=%d  r2hc %.3e  ord %.3e  hc2r(ord) %.3e  hc2r(hc) %.3e  %s\n", sizeof(T) == 4 ? "f32" : "f64",
           LOG2MS, R0, e1, e2, e3[0], e3[1], ok ? "ok" : "FAIL");
    if (!ok) fails++;
    return fails;
}

int main()
{
    int f = 0;
    f += check<float, 4, 1>(2e-6); f += check<float, 5, 1>(2e-6); f += check<float, 6, 1>(2e-6); f += check<float, 7, 1>(2e-6);
    f += check<float, 8, 1>(2e-6); f += check<float, 9, 1>(2e-6); f += check<float, 10, 1>(2e-6); f += check<float, 11, 1>(2e-6);
    f += check<float, 12, 1>(2e-6); f += check<float, 13, 1>(2e-6); f += check<float, 14, 1>(2e-6);
    f += check<double, 4, 1>(4e-15); f += check<double, 5, 1>(4e-15); f += check<double, 6, 1>(4e-15); f += check<double, 7, 1>(4e-15);
    f += check<double, 8, 1>(4e-15); f += check<double, 9, 1>(4e-15); f += check<double, 10, 1>(4e-15); f += check<double, 11, 1>(4e-15);
    f += check<double, 12, 1>(4e-15); f += check<double, 13, 1>(4e-15);
    f += check<float, 4, 2>(2e-6); f += check<float, 7, 2>(2e-6); f += check<float, 9, 2>(2e-6); f += check<float, 12, 2>(2e-6);
    f += check<float, 14, 2>(2e-6);
    f += check<double, 5, 2>(4e-15); f += check<double, 8, 2>(4e-15); f += check<double, 12, 2>(4e-15); f += check<double, 13, 2>(4e-15);
    printf(f ? "FAILED %d\n" : "ALL OK\n", f);
    return f ? 1 : 0;
}
