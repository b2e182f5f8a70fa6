// CPU emulation of the CTA-wide real FFT kernels: every phase of rfft_forward_kernel /
// rfft_inverse_kernel is executed for all threads in turn (a __syncthreads boundary = end of a loop
// over t), using the very same __host__ __device__ functions the GPU runs. Checked against the
// oracle FFT (oracle/fft_r2r.hpp). Runs without a GPU; used by tests/test_host_emulation.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <cstring>
#include <cstdint>
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
template <class T, int LOG2MS, int R0, int LOG2E = 4> static int check(double tol)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, M = MS * R0, N = 2 * M, NT = MS / E;
    typedef cpx<T> C;
    std::mt19937 rng(1234 + LOG2MS + 100 * R0);
    std::uniform_real_distribution<double> u(-1, 1);
    std::vector<T> x(N), hc(N), ord(N), ref(N), back(N);
    for (auto &v : x) v = (T)u(rng);
    auto tw = make_tw<T>(N);
    const int sm = R0 == 2 ? 2 : 1; // log2(N / MS)
    std::vector<C> smem(fft_smem_elems<MS>::value);
    std::vector<C> regs((size_t)NT * E);
    C (*vs)[E] = reinterpret_cast<C (*)[E]>(regs.data());
    int fails = 0;

    for (int layout = 0; layout < 2; layout++) {
        FwdArgs a = {};
        a.in_mode = IN_TIME; a.out_layout = layout; a.in = x.data(); a.out = layout == LAYOUT_HC ? hc.data() : ord.data();
        a.scale_in = 1.0; a.scale_out = layout == LAYOUT_ORD ? 0.5 : 1.0;
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) fwd_load<T, LOG2MS, R0, LOG2E>(t, 0, 0, r, vs[t], tw.data(), 0, a);
            fft_passes<T, LOG2MS, false, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) BlockFFT<T, LOG2MS, false, LOG2E>::store_natural(t, vs[t], smem.data());
            for (int t = 0; t < NT; t++) fwd_split_store<T, LOG2MS, R0, LOG2E>(t, 0, 0, r, smem.data(), tw.data(), 0, a);
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
            for (int t = 0; t < NT; t++) inv_load<T, LOG2MS, R0, LOG2E>(t, 0, r, vs[t], tw.data(), 0, b);
            fft_passes<T, LOG2MS, true, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            OverflowAcc acc = {};
            for (int t = 0; t < NT; t++) inv_store<T, LOG2MS, R0, LOG2E>(t, 0, r, vs[t], b, acc);
        }
        std::vector<T> want(N);
        for (int i = 0; i < N; i++) want[i] = x[i] * (T)N;
        e3[layout] = rel_rms(back, want);
    }
    bool ok = e1 < tol && e2 < tol && e3[0] < tol && e3[1] < tol;
    printf("%s log2ms=%2d R0=%d E=%2d  r2hc %.3e  ord %.3e  hc2r(ord) %.3e  hc2r(hc) %.3e  %s\n", sizeof(T) == 4 ? "f32" : "f64",
           LOG2MS, R0, E, e1, e2, e3[0], e3[1], ok ? "ok" : "FAIL");
    if (!ok) fails++;
    return fails;
}

// Engine modes of the same kernels (raw interleaved input with the previous-block ping-pong, delay-line
// slot addressing, coefficient partitions, raw / planar output) on exactly-sized host buffers, so that an
// out-of-bounds index shows up under AddressSanitizer. Values are checked against the IN_TIME/OUT_TIME
// path verified above.
template <class T, int LOG2MS, int R0, int LOG2E = 4> static int check_engine_modes(int fmt_in, int fmt_out)
{
    constexpr int E = 1 << LOG2E, MS = 1 << LOG2MS, M = MS * R0, N = 2 * M, L = M, NT = MS / E;
    typedef cpx<T> C;
    const int S = 2, CH = 3, Ct = S * CH, P = 3;
    const int sm = R0 == 2 ? 2 : 1;
    std::mt19937 rng(77);
    std::uniform_real_distribution<double> u(-1, 1);
    auto tw = make_tw<T>(N);
    std::vector<C> smem(fft_smem_elems<MS>::value);
    std::vector<C> regs((size_t)NT * E);
    C (*vs)[E] = reinterpret_cast<C (*)[E]>(regs.data());
    const int ib = fmt_bytes(fmt_in), ob = fmt_bytes(fmt_out);
    std::vector<uint8_t> raw_in((size_t)S * L * CH * ib), raw_out((size_t)S * L * CH * ob, 0);
    std::vector<T> planar((size_t)Ct * L);
    for (int s = 0; s < S; s++) for (int n = 0; n < L; n++) for (int c = 0; c < CH; c++) {
        const float v = (float)u(rng);
        uint8_t *p = &raw_in[(((size_t)s * L + n) * CH + c) * ib];
        if (fmt_in == FMT_FLOAT_LE) memcpy(p, &v, 4);
        else { const int32_t q = (int32_t)lrint(v * 8388607.0); p[0] = q & 0xff; p[1] = (q >> 8) & 0xff; p[2] = (q >> 16) & 0xff; }
        planar[((size_t)s * CH + c) * L + n] = load_raw<T>(p, fmt_in);
    }
    std::vector<T> prev((size_t)2 * Ct * L, (T)0), fdl((size_t)Ct * P * N, (T)0), ref(N), tbuf(N);
    std::vector<int> procblocks(Ct, 0);
    std::vector<unsigned char> pb_inc(Ct, 0);
    EngineState st; st.blockcounter = 4; st.first_bad_channel = 0x7fffffff;   // slot 4 % 3 = 1, parity 0
    int fails = 0;
    for (int bx = 0; bx < Ct; bx++) {
        FwdArgs a = {};
        a.in_mode = IN_RAW_PREV; a.out_layout = LAYOUT_ORD; a.in = raw_in.data(); a.in_stride_x = (long long)L * CH * ib;
        a.out = fdl.data(); a.out_stride_x = (long long)P * N; a.out_stride_y = N; a.scale_in = 1.0; a.scale_out = 0.5;
        a.prev = prev.data(); a.fmt = fmt_in; a.ch_per_stream = CH; a.n_channels = Ct; a.ch_base = 0; a.prev_parity = 0;
        a.state = &st; a.n_slots = P; a.n_parts = P; a.procblocks = procblocks.data(); a.pb_inc = pb_inc.data();
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) fwd_load<T, LOG2MS, R0, LOG2E>(t, bx, 0, r, vs[t], tw.data(), 0, a);
            fft_passes<T, LOG2MS, false, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) BlockFFT<T, LOG2MS, false, LOG2E>::store_natural(t, vs[t], smem.data());
            for (int t = 0; t < NT; t++) fwd_split_store<T, LOG2MS, R0, LOG2E>(t, bx, 0, r, smem.data(), tw.data(), 0, a);
        }
        // reference: [0 | cur] through the plain path
        std::fill(tbuf.begin(), tbuf.end(), (T)0);
        for (int n = 0; n < L; n++) tbuf[L + n] = planar[(size_t)bx * L + n];
        FwdArgs b = {};
        b.in_mode = IN_TIME; b.out_layout = LAYOUT_ORD; b.in = tbuf.data(); b.out = ref.data(); b.scale_in = 1.0; b.scale_out = 0.5;
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) fwd_load<T, LOG2MS, R0, LOG2E>(t, 0, 0, r, vs[t], tw.data(), 0, b);
            fft_passes<T, LOG2MS, false, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) BlockFFT<T, LOG2MS, false, LOG2E>::store_natural(t, vs[t], smem.data());
            for (int t = 0; t < NT; t++) fwd_split_store<T, LOG2MS, R0, LOG2E>(t, 0, 0, r, smem.data(), tw.data(), 0, b);
        }
        std::vector<T> got(fdl.begin() + ((size_t)bx * P + 1) * N, fdl.begin() + ((size_t)bx * P + 2) * N);
        if (rel_rms(got, ref) > 1e-6) { printf("engine fwd mismatch ch %d\n", bx); fails++; }
        // the current block must now sit in the OTHER half of the ping-pong pair
        for (int n = 0; n < L; n++) if (prev[((size_t)1 * Ct + bx) * L + n] != planar[(size_t)bx * L + n]) { fails++; break; }
        if (procblocks[bx] != 1 || pb_inc[bx] != 1) fails++;
    }
    // inverse: delay-line slot -> raw interleaved output, must reproduce [0 | cur] * N * 0.5 * 2 in the first half... use scale 1/N
    std::vector<OverflowStats> stats(Ct);
    for (int bx = 0; bx < Ct; bx++) {
        InvArgs v = {};
        v.in_layout = LAYOUT_ORD; v.in = fdl.data() + (size_t)1 * N; v.in_stride_x = (long long)P * N; v.scale_in = 2.0 / N;
        v.out_mode = OUT_RAW; v.out = raw_out.data(); v.out_stride_x = (long long)L * CH * ob; v.fmt = fmt_out; v.ch_per_stream = CH;
        v.ovf_max = fmt_isfloat(fmt_out) ? 1.0 : 32767.0; v.stats = stats.data(); v.state = &st;
        if (!fmt_isfloat(fmt_out)) v.scale_in *= 32768.0 / (fmt_in == FMT_FLOAT_LE ? 1.0 : 8388608.0);
        OverflowAcc acc = {};
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) inv_load<T, LOG2MS, R0, LOG2E>(t, bx, r, vs[t], tw.data(), 0, v);
            fft_passes<T, LOG2MS, true, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) inv_store<T, LOG2MS, R0, LOG2E>(t, bx, r, vs[t], v, acc);
        }
    }
    // [0 | cur] -> first half of the inverse is zero (the previous block was silence)
    double worst = 0;
    for (size_t i = 0; i < raw_out.size() / ob; i++) {
        double val;
        if (fmt_out == FMT_FLOAT_LE) { float f; memcpy(&f, &raw_out[i * 4], 4); val = f; }
        else { int16_t q; memcpy(&q, &raw_out[i * 2], 2); val = q / 32768.0; }
        worst = std::fmax(worst, std::fabs(val));
    }
    const double amp = fmt_in == FMT_FLOAT_LE ? 1.0 : 8388608.0;   // integer input is not normalised in this check
    if (worst > (fmt_out == FMT_FLOAT_LE ? 1e-5 * amp : 2.0 / 32768)) { printf("engine inverse: expected silence, got %g\n", worst); fails++; }
    {   // look-ahead mode: `in` = partitions 1..P-1 (zeros here), head term X[slot] * H[0] added in the load phase.
        // H[0] = 1 + 0i for every bin (ORD: real parts 1, imaginary parts 0, Nyquist in slot 4 = 1) must reproduce
        // the plain inverse of the slot; a channel without coefficients (head_blocks 0) must stay silent.
        std::vector<T> zacc((size_t)Ct * N, (T)0), hone((size_t)Ct * P * N, (T)0);
        for (int c = 0; c < Ct; c++) for (int g8 = 0; g8 < N / 8; g8++) for (int j = 0; j < 4; j++) hone[((size_t)c * P) * N + g8 * 8 + j] = (T)1;
        for (int c = 0; c < Ct; c++) hone[((size_t)c * P) * N + 4] = (T)1;
        std::vector<int> hb(Ct, 1);
        hb[Ct - 1] = 0;
        std::vector<uint8_t> raw_head((size_t)S * L * CH * ob, 0xee);
        if (st.cur_slot != 1u) { printf("cur_slot %u\n", st.cur_slot); fails++; }
        for (int bx = 0; bx < Ct; bx++) {
            InvArgs v = {};
            v.in_layout = LAYOUT_ORD; v.in = zacc.data(); v.in_stride_x = N; v.scale_in = 2.0 / N;
            v.out_mode = OUT_RAW; v.out = raw_head.data(); v.out_stride_x = (long long)L * CH * ob; v.fmt = fmt_out; v.ch_per_stream = CH;
            v.ovf_max = fmt_isfloat(fmt_out) ? 1.0 : 32767.0; v.stats = stats.data(); v.state = &st;
            if (!fmt_isfloat(fmt_out)) v.scale_in *= 32768.0 / (fmt_in == FMT_FLOAT_LE ? 1.0 : 8388608.0);
            v.head_x = fdl.data(); v.head_x_stride = (long long)P * N; v.head_h = hone.data(); v.head_h_stride = (long long)P * N;
            v.head_blocks = hb.data();
            OverflowAcc acc = {};
            for (int r = 0; r < R0; r++) {
                for (int t = 0; t < NT; t++) inv_load<T, LOG2MS, R0, LOG2E>(t, bx, r, vs[t], tw.data(), 0, v);
                fft_passes<T, LOG2MS, true, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
                for (int t = 0; t < NT; t++) inv_store<T, LOG2MS, R0, LOG2E>(t, bx, r, vs[t], v, acc);
            }
        }
        for (int s_ = 0; s_ < S; s_++) for (int n = 0; n < L; n++) for (int c = 0; c < CH; c++) {
            const size_t o = (((size_t)s_ * L + n) * CH + c) * ob;
            const bool silent = s_ * CH + c == Ct - 1;
            for (int k = 0; k < ob; k++) {
                const uint8_t want = silent ? 0 : raw_out[o + k];
                if (raw_head[o + k] != want && !(silent && fmt_isfloat(fmt_out) && k == ob - 1 && raw_head[o + k] == 0x80)) { // -0.0f
                    if (fails < 5) printf("head mode mismatch s %d n %d c %d\n", s_, n, c);
                    fails++; break;
                }
            }
        }
    }
    // coefficient partitions (IN_COEFF) and the upper-half mode (IN_UPPER)
    std::vector<T> h(2 * L + 5), hs((size_t)P * N, (T)7), up(N);
    for (auto &x : h) x = (T)u(rng);
    int flag = 0;
    for (int by = 0; by < P; by++) {
        FwdArgs a = {};
        a.in_mode = IN_COEFF; a.out_layout = LAYOUT_ORD; a.in = h.data(); a.in_stride_x = (long long)h.size();
        a.out = hs.data(); a.out_stride_x = (long long)P * N; a.out_stride_y = N; a.scale_in = 0.5; a.scale_out = 1.0 / N;
        a.coeff_len = (int)h.size(); a.nonfinite = &flag;
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) fwd_load<T, LOG2MS, R0, LOG2E>(t, 0, by, r, vs[t], tw.data(), 0, a);
            fft_passes<T, LOG2MS, false, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) BlockFFT<T, LOG2MS, false, LOG2E>::store_natural(t, vs[t], smem.data());
            for (int t = 0; t < NT; t++) fwd_split_store<T, LOG2MS, R0, LOG2E>(t, 0, by, r, smem.data(), tw.data(), 0, a);
        }
    }
    if (flag != 0) fails++;
    {   // partition 2 holds only 5 coefficients: compare with IN_UPPER on the same padded data
        std::vector<T> src(L, (T)0);
        for (int n = 0; n < 5; n++) src[n] = h[2 * L + n] * (T)0.5;
        FwdArgs a = {};
        a.in_mode = IN_UPPER; a.out_layout = LAYOUT_ORD; a.in = src.data(); a.out = up.data(); a.scale_in = 1.0; a.scale_out = 1.0 / N;
        for (int r = 0; r < R0; r++) {
            for (int t = 0; t < NT; t++) fwd_load<T, LOG2MS, R0, LOG2E>(t, 0, 0, r, vs[t], tw.data(), 0, a);
            fft_passes<T, LOG2MS, false, 0, 0, LOG2E>::run_host(vs, smem.data(), tw.data(), sm);
            for (int t = 0; t < NT; t++) BlockFFT<T, LOG2MS, false, LOG2E>::store_natural(t, vs[t], smem.data());
            for (int t = 0; t < NT; t++) fwd_split_store<T, LOG2MS, R0, LOG2E>(t, 0, 0, r, smem.data(), tw.data(), 0, a);
        }
        std::vector<T> got(hs.begin() + (size_t)2 * N, hs.end());
        if (rel_rms(got, up) > 1e-6) { printf("coefficient partition mismatch\n"); fails++; }
    }
    printf("%s engine modes log2ms=%d R0=%d E=%d fmt %d->%d  %s\n", sizeof(T) == 4 ? "f32" : "f64", LOG2MS, R0, E, fmt_in, fmt_out, fails ? "FAIL" : "ok");
    return fails;
}

int main(int argc, char **argv)
{
    const bool quick = argc > 1 && strcmp(argv[1], "quick") == 0;   // small sizes only (the ASan run of the test suite)
    int f = 0;
    f += check_engine_modes<float, 6, 1>(FMT_FLOAT_LE, FMT_FLOAT_LE);
    f += check_engine_modes<double, 5, 1>(FMT_S24_LE, FMT_S16_LE);
    f += check_engine_modes<float, 5, 2>(FMT_S24_LE, FMT_FLOAT_LE);
    f += check_engine_modes<double, 7, 2>(FMT_FLOAT_LE, FMT_S16_LE);
    // every pass plan shape once: [16], [8,4], [16,4], [16,8], [16,16], [16,8,4]
    f += check<float, 4, 1>(2e-6); f += check<double, 5, 1>(4e-15); f += check<float, 6, 1>(2e-6); f += check<double, 7, 1>(4e-15);
    f += check<float, 8, 1>(2e-6); f += check<double, 9, 1>(4e-15);
    f += check<float, 7, 2>(2e-6); f += check<double, 5, 2>(4e-15);
    // 8 points per thread (the double-precision variants): pass plans [8,8], [8,8,2], [8,8,4], one and two CTAs, engine modes
    f += check<double, 6, 1, 3>(4e-15); f += check<double, 7, 1, 3>(4e-15); f += check<double, 8, 1, 3>(4e-15);
    f += check<double, 6, 2, 3>(4e-15); f += check<double, 7, 2, 3>(4e-15);
    f += check_engine_modes<double, 6, 1, 3>(FMT_S24_LE, FMT_S16_LE);
    f += check_engine_modes<double, 7, 2, 3>(FMT_FLOAT_LE, FMT_FLOAT_LE);
#ifndef EMU_QUICK
    if (!quick) {
        f += check<float, 5, 1>(2e-6); f += check<float, 7, 1>(2e-6); f += check<float, 9, 1>(2e-6); f += check<float, 10, 1>(2e-6);
        f += check<double, 4, 1>(4e-15); f += check<double, 6, 1>(4e-15); f += check<double, 8, 1>(4e-15); f += check<double, 10, 1>(4e-15);
        f += check<float, 4, 2>(2e-6); f += check<float, 9, 2>(2e-6); f += check<double, 8, 2>(4e-15);
        f += check<float, 11, 1>(2e-6); f += check<float, 12, 1>(2e-6); f += check<float, 13, 1>(2e-6); f += check<float, 14, 1>(2e-6);
        f += check<double, 11, 1>(4e-15); f += check<double, 12, 1>(4e-15); f += check<double, 13, 1>(4e-15);
        f += check<float, 12, 2>(2e-6); f += check<float, 14, 2>(2e-6);
        f += check<double, 12, 2>(4e-15); f += check<double, 13, 2>(4e-15);
        f += check<double, 9, 1, 3>(4e-15); f += check<double, 10, 1, 3>(4e-15); f += check<double, 11, 1, 3>(4e-15); f += check<double, 12, 1, 3>(4e-15);
        f += check<double, 9, 2, 3>(4e-15); f += check<double, 10, 2, 3>(4e-15); f += check<double, 11, 2, 3>(4e-15); f += check<double, 12, 2, 3>(4e-15);
    }
#endif
    printf(f ? "FAILED %d\n" : "ALL OK\n", f);
    return f ? 1 : 0;
}
