// Block-wide complex FFT held in registers + one shared-memory buffer (sm_100a).
//
// Replaces the FFTW plans of the reference (fftw[f]_plan_r2r_1d / fftw[f]_execute_r2r, reference
// brutefir/fftw_convolver.cpp:188-212, 351-375, 780-817): an N-point real transform (N = 2 * block
// length) is one M = N/2 point complex transform plus a split step, all inside one CTA.
//
// Geometry: M = 2^LOG2M points, NT = M/16 threads, every thread owns E = 16 points. The transform is
// a decimation-in-frequency Stockham autosort with radix 16/8/4/2 passes:
//     pass with radix R, stride s (product of earlier radices), butterfly b in [0, M/R):
//         p = b / s, q = b % s
//         inputs   x[b + k*M/R]            k = 0..R-1
//         outputs  y[q + s*(R*p + k)] = DFT_R(inputs)[k] * W_M^(p*s*k)
// With 16 points per thread, butterfly j of thread t is b = t + j*NT and its inputs are exactly the
// thread's registers v[j + k*(16/R)] = x[t + (j + k*16/R)*NT]: every pass reads the SAME 16 shared
// memory addresses t + i*NT (conflict free) and only the store is permuted. After the last pass the
// registers already hold the result in natural order (v[i] = X[t + i*NT]), so the first pass loads
// straight from global memory and the last pass stores straight to global memory.
//
// Everything that touches data is __host__ __device__ so the whole block can be emulated thread by
// thread on the CPU (tests/host_emulation), where no GPU exists.
#pragma once
#include <cuda_runtime.h>

#ifndef BFIR_HD
#define BFIR_HD __host__ __device__ __forceinline__
#endif

namespace bfir {

template <class T> struct cpx_of;
template <> struct cpx_of<float> { typedef float2 type; };
template <> struct cpx_of<double> { typedef double2 type; };
template <class T> using cpx = typename cpx_of<T>::type;

template <class T> BFIR_HD cpx<T> mk(T x, T y) { cpx<T> r; r.x = x; r.y = y; return r; }
template <class C> BFIR_HD C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <class C> BFIR_HD C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <class C> BFIR_HD C cmul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <class C> BFIR_HD C cconj(C a) { a.y = -a.y; return a; }
// multiply by -i (forward) or +i (inverse)
template <bool INV, class C> BFIR_HD C mul_mi(C a) { C r; if (INV) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; } return r; }

// ---------------------------------------------------------------------------------------------
// small DFTs on registers, natural order in and out. Forward: W = exp(-2 pi i / R); INV conjugates.
template <bool INV, class C> BFIR_HD void dft2(C &a, C &b) { C t = a; a = cadd(t, b); b = csub(t, b); }

template <bool INV, class C> BFIR_HD void dft4(C &a0, C &a1, C &a2, C &a3)
{
    C s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a1 = cadd(d02, d13);
    a2 = csub(s02, s13);
    a3 = csub(d02, d13);
}

// multiply by W8^1 = (1 - i)/sqrt2 (fwd) and W8^3 = (-1 - i)/sqrt2 (fwd)
template <bool INV, class T> BFIR_HD cpx<T> mul_w8_1(cpx<T> a)
{
    const T h = (T)0.70710678118654752440;
    cpx<T> r;
    if (INV) { r.x = (a.x - a.y) * h; r.y = (a.x + a.y) * h; } else { r.x = (a.x + a.y) * h; r.y = (a.y - a.x) * h; }
    return r;
}
template <bool INV, class T> BFIR_HD cpx<T> mul_w8_3(cpx<T> a)
{
    const T h = (T)0.70710678118654752440;
    cpx<T> r;
    if (INV) { r.x = (-a.x - a.y) * h; r.y = (a.x - a.y) * h; } else { r.x = (a.y - a.x) * h; r.y = (-a.x - a.y) * h; }
    return r;
}

// n = n1 + 2 n2, k = 4 k1 + k2:  W8^(nk) = W2^(n1 k1) W8^(n1 k2) W4^(n2 k2)
template <bool INV, class T> BFIR_HD void dft8(cpx<T> (&a)[8])
{
    dft4<INV>(a[0], a[2], a[4], a[6]); // n1 = 0 -> A[0][k2] in a[0],a[2],a[4],a[6]
    dft4<INV>(a[1], a[3], a[5], a[7]); // n1 = 1 -> A[1][k2] in a[1],a[3],a[5],a[7]
    a[3] = mul_w8_1<INV, T>(a[3]);
    a[5] = mul_mi<INV>(a[5]);
    a[7] = mul_w8_3<INV, T>(a[7]);
    // X[4 k1 + k2] = A[0][k2] + (-1)^k1 A[1][k2]
    cpx<T> x0 = cadd(a[0], a[1]), x4 = csub(a[0], a[1]);
    cpx<T> x1 = cadd(a[2], a[3]), x5 = csub(a[2], a[3]);
    cpx<T> x2 = cadd(a[4], a[5]), x6 = csub(a[4], a[5]);
    cpx<T> x3 = cadd(a[6], a[7]), x7 = csub(a[6], a[7]);
    a[0] = x0; a[1] = x1; a[2] = x2; a[3] = x3; a[4] = x4; a[5] = x5; a[6] = x6; a[7] = x7;
}

// multiply by W16^j (forward) for the j that occur in the 4x4 decomposition
template <bool INV, int J, class T> BFIR_HD cpx<T> mul_w16(cpx<T> a)
{
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    // forward twiddle W16^J = (wr, wi) with wi <= 0 pattern; INV uses the conjugate
    T wr, wi;
    if (J == 0) return a;
    if (J == 1) { wr = c1; wi = -s1; }
    else if (J == 2) { wr = h; wi = -h; }
    else if (J == 3) { wr = s1; wi = -c1; }
    else if (J == 4) { return mul_mi<INV>(a); }
    else if (J == 6) { wr = -h; wi = -h; }
    else /* J == 9 */ { wr = -c1; wi = s1; }
    if (INV) wi = -wi;
    cpx<T> r;
    r.x = a.x * wr - a.y * wi;
    r.y = a.x * wi + a.y * wr;
    return r;
}

// n = n1 + 4 n2, k = 4 k1 + k2:  W16^(nk) = W4^(n1 k1) W16^(n1 k2) W4^(n2 k2)
template <bool INV, class T> BFIR_HD void dft16(cpx<T> (&a)[16])
{
    // step 1: DFT4 over n2 for each n1; result A[n1][k2] stays in a[n1 + 4 k2]
    dft4<INV>(a[0], a[4], a[8], a[12]);
    dft4<INV>(a[1], a[5], a[9], a[13]);
    dft4<INV>(a[2], a[6], a[10], a[14]);
    dft4<INV>(a[3], a[7], a[11], a[15]);
    // twiddle A[n1][k2] *= W16^(n1 k2)
    a[5] = mul_w16<INV, 1, T>(a[5]);   a[9] = mul_w16<INV, 2, T>(a[9]);   a[13] = mul_w16<INV, 3, T>(a[13]);
    a[6] = mul_w16<INV, 2, T>(a[6]);   a[10] = mul_w16<INV, 4, T>(a[10]); a[14] = mul_w16<INV, 6, T>(a[14]);
    a[7] = mul_w16<INV, 3, T>(a[7]);   a[11] = mul_w16<INV, 6, T>(a[11]); a[15] = mul_w16<INV, 9, T>(a[15]);
    // step 2: DFT4 over n1 for each k2: inputs a[0 + 4k2], a[1 + 4k2], a[2 + 4k2], a[3 + 4k2] -> X[4 k1 + k2]
    dft4<INV>(a[0], a[1], a[2], a[3]);     // k2 = 0 -> X[0], X[4], X[8], X[12]
    dft4<INV>(a[4], a[5], a[6], a[7]);     // k2 = 1 -> X[1], X[5], X[9], X[13]
    dft4<INV>(a[8], a[9], a[10], a[11]);   // k2 = 2 -> X[2], X[6], X[10], X[14]
    dft4<INV>(a[12], a[13], a[14], a[15]); // k2 = 3 -> X[3], X[7], X[11], X[15]
    // a[4 k2 + k1] holds X[4 k1 + k2]: transpose the 4x4 index
    cpx<T> t;
    t = a[1]; a[1] = a[4]; a[4] = t;
    t = a[2]; a[2] = a[8]; a[8] = t;
    t = a[3]; a[3] = a[12]; a[12] = t;
    t = a[6]; a[6] = a[9]; a[9] = t;
    t = a[7]; a[7] = a[13]; a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}

template <int R, bool INV, class T> struct small_dft;
template <bool INV, class T> struct small_dft<2, INV, T> { static BFIR_HD void run(cpx<T> (&a)[2]) { dft2<INV>(a[0], a[1]); } };
template <bool INV, class T> struct small_dft<4, INV, T> { static BFIR_HD void run(cpx<T> (&a)[4]) { dft4<INV>(a[0], a[1], a[2], a[3]); } };
template <bool INV, class T> struct small_dft<8, INV, T> { static BFIR_HD void run(cpx<T> (&a)[8]) { dft8<INV, T>(a); } };
template <bool INV, class T> struct small_dft<16, INV, T> { static BFIR_HD void run(cpx<T> (&a)[16]) { dft16<INV, T>(a); } };

// exp(-2 pi i j / D) for j = 0..15 and D = 32 or 64, as compile-time constants: the 16 points a thread
// owns are N/32 (or N/64) bins apart, so W_N^(k0 + j N/D) = W_N^k0 * root<D>(j) needs ONE table look-up
// per thread instead of 16
template <class T, int D> BFIR_HD cpx<T> unit_root(int j)
{
    constexpr double c32[16] = { 1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
        0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785,
        0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
        -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913 };
    constexpr double s32[16] = { 0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
        0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913,
        1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
        0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785 };
    constexpr double c64[16] = { 1.0, 0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
        0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708, 0.77301045336273696081,
        0.70710678118654752440, 0.63439328416364549822, 0.55557023301960222474, 0.47139673682599764856,
        0.38268343236508977173, 0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199 };
    constexpr double s64[16] = { 0.0, 0.09801714032956060199, 0.19509032201612826785, 0.29028467725446236764,
        0.38268343236508977173, 0.47139673682599764856, 0.55557023301960222474, 0.63439328416364549822,
        0.70710678118654752440, 0.77301045336273696081, 0.83146961230254523708, 0.88192126434835502971,
        0.92387953251128675613, 0.95694033573220886494, 0.98078528040323044913, 0.99518472667219688624 };
    return D == 32 ? mk<T>((T)c32[j], (T)-s32[j]) : mk<T>((T)c64[j], (T)-s64[j]);
}

// root of unity between consecutive points of one thread: D = 2E (or 4E with two CTAs per transform)
template <class T, int D> BFIR_HD cpx<T> thread_root(int j)
{
    static_assert(D == 16 || D == 32 || D == 64, "8 or 16 points per thread");
    if (D == 16) return unit_root<T, 32>(2 * j);      // j < 8
    return unit_root<T, D == 16 ? 32 : D>(j);
}

// ---------------------------------------------------------------------------------------------
// pass plan: radices whose product is 2^LOG2M, none above the E = 2^LOG2E points a thread owns, largest first
// (E = 16: 16s first, so strides are >= 16 from pass 2 on; a remainder of 2^5 is split 8 x 4 rather than 16 x 2)
template <int LOG2M, int LOG2E = 4> struct fft_plan {
    static constexpr int n16 = LOG2E == 4 ? ((LOG2M % 4 == 1 && LOG2M >= 5) ? LOG2M / 4 - 1 : LOG2M / 4) : LOG2M / LOG2E;
    static constexpr int rem = LOG2M - LOG2E * n16;         // E = 16: 0,1,2,3 or 5 (-> 8*4); E = 8: 0,1,2
    static constexpr int npass = n16 + (rem == 0 ? 0 : (rem == 5 ? 2 : 1));
    static constexpr BFIR_HD int log2_radix(int pass)
    {
        return pass < n16 ? LOG2E : (rem == 5 ? (pass == n16 ? 3 : 2) : rem);
    }
};

// shared-memory index padding: one extra element per 16 keeps the radix-16 first-pass store (stride 16
// elements between threads) off a single bank
BFIR_HD int fft_pad(int i) { return i + (i >> 4); }
template <int M> struct fft_smem_elems { static constexpr int value = M + (M >> 4); };

// LOG2E = 4: 16 points per thread (the default everywhere); LOG2E = 3: 8 points per thread on twice the threads,
// i.e. half the dependent work per thread and twice the warps per transform (double precision, where 16 points
// cost 128 registers and leave two warps per scheduler)
template <class T, int LOG2M, bool INV, int LOG2E = 4>
struct BlockFFT {
    static constexpr int M = 1 << LOG2M;
    static_assert(LOG2M >= 4, "block length below 16 is not supported");
    static_assert(LOG2E == 3 || LOG2E == 4, "8 or 16 points per thread");
    static constexpr int E = 1 << LOG2E;               // points per thread
    static constexpr int NT = M / E;                   // threads
    static constexpr int LOG2NT = LOG2M - LOG2E;
    typedef fft_plan<LOG2M, LOG2E> plan;
    typedef cpx<T> C;

    // table look-ups of one pass for thread t: per butterfly w^1, and w^4 for radix 8/16 (the other powers
    // follow by at most 3 multiplications). Kept apart from the butterflies so that the loads can be issued
    // one pass ahead, before the shared-memory exchange, and their latency stays off the critical path.
    //   tw: table of exp(-2 pi i j / NTW), tw_shift = log2(NTW / M)
    template <int LOG2R> static constexpr int tw_regs() { return (E >> LOG2R) * (LOG2R >= 3 ? 2 : 1); }

    template <int LOG2R, int LOG2S>
    static BFIR_HD void load_twiddles(int t, C (&w)[tw_regs<LOG2R>()], const C *__restrict__ tw, int tw_shift)
    {
        constexpr int NB = E >> LOG2R, PER = LOG2R >= 3 ? 2 : 1;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const int b = t + j * NT;
            const int ps = (b >> LOG2S) << LOG2S; // p * s
            w[PER * j] = tw[ps << tw_shift];
            if constexpr (PER == 2) w[PER * j + 1] = tw[(ps * 4) << tw_shift];
        }
    }

    // Butterflies + twiddles of one pass on the thread's registers.
    template <int LOG2R, int LOG2S, bool LAST>
    static BFIR_HD void butterflies(int t, C (&v)[E], const C (&w)[tw_regs<LOG2R>()])
    {
        constexpr int R = 1 << LOG2R;
        constexpr int NB = E / R; // butterflies per thread
        constexpr int PER = LOG2R >= 3 ? 2 : 1;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            C a[R];
#pragma unroll
            for (int k = 0; k < R; k++) a[k] = v[j + k * NB];
            small_dft<R, INV, T>::run(a);
            if constexpr (!LAST) {
                C w1 = w[PER * j];
                if (INV) w1 = cconj(w1);
                if constexpr (R == 2) {
                    a[1] = cmul(a[1], w1);
                } else if constexpr (R == 4) {
                    C w2 = cmul(w1, w1);
                    a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], cmul(w2, w1));
                } else {
                    C w4 = w[PER * j + 1];
                    if (INV) w4 = cconj(w4);
                    C w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                    a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], w3);
                    a[4] = cmul(a[4], w4); a[5] = cmul(a[5], cmul(w4, w1)); a[6] = cmul(a[6], cmul(w4, w2)); a[7] = cmul(a[7], cmul(w4, w3));
                    if constexpr (R == 16) {
                        C w8 = cmul(w4, w4), w12 = cmul(w8, w4);
                        a[8] = cmul(a[8], w8); a[9] = cmul(a[9], cmul(w8, w1));
                        a[10] = cmul(a[10], cmul(w8, w2)); a[11] = cmul(a[11], cmul(w8, w3));
                        a[12] = cmul(a[12], w12); a[13] = cmul(a[13], cmul(w12, w1));
                        a[14] = cmul(a[14], cmul(w12, w2)); a[15] = cmul(a[15], cmul(w12, w3));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < R; k++) v[j + k * NB] = a[k];
        }
    }

    // permuted store of a non-final pass: y[q + s*(R*p + k)]
    template <int LOG2R, int LOG2S>
    static BFIR_HD void store_pass(int t, const C (&v)[E], C *smem)
    {
        constexpr int R = 1 << LOG2R;
        constexpr int NB = E / R;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const int b = t + j * NT;
            const int q = b & ((1 << LOG2S) - 1);
            const int p = b >> LOG2S;
            const int base = q + (p << (LOG2S + LOG2R));
#pragma unroll
            for (int k = 0; k < R; k++) smem[fft_pad(base + (k << LOG2S))] = v[j + k * NB];
        }
    }

    static BFIR_HD void load_natural(int t, C (&v)[E], const C *smem)
    {
#pragma unroll
        for (int i = 0; i < E; i++) v[i] = smem[fft_pad(t + i * NT)];
    }

    static BFIR_HD void store_natural(int t, const C (&v)[E], C *smem)
    {
#pragma unroll
        for (int i = 0; i < E; i++) smem[fft_pad(t + i * NT)] = v[i];
    }
};

// ---------------------------------------------------------------------------------------------
// The pass loop, written once with a SYNC functor so the device kernel passes __syncthreads and the
// host emulation runs each phase for all threads in turn. PASS is unrolled by recursion because the
// radix and stride of every pass are compile-time constants.
template <class T, int LOG2M, bool INV, int PASS, int LOG2S, int LOG2E = 4>
struct fft_passes {
    typedef BlockFFT<T, LOG2M, INV, LOG2E> F;
    typedef fft_plan<LOG2M, LOG2E> plan;
    static constexpr int LOG2R = plan::log2_radix(PASS);
    static constexpr bool LAST = (PASS == plan::npass - 1);

    static constexpr int NW = F::template tw_regs<LOG2R>();

#ifdef __CUDACC__
    // device: all threads of the CTA call run(); on return v[i] = X[t + i*NT] (natural order)
    static __device__ __forceinline__ void run(int t, cpx<T> (&v)[F::E], cpx<T> *smem, const cpx<T> *__restrict__ tw, int tw_shift)
    {
        static_assert(PASS == 0, "enter at the first pass");
        cpx<T> w[NW];
        if constexpr (!LAST) F::template load_twiddles<LOG2R, LOG2S>(t, w, tw, tw_shift);
        pass(t, v, smem, tw, tw_shift, w);
    }

    static __device__ __forceinline__ void pass(int t, cpx<T> (&v)[F::E], cpx<T> *smem, const cpx<T> *__restrict__ tw, int tw_shift, const cpx<T> (&w)[NW])
    {
        F::template butterflies<LOG2R, LOG2S, LAST>(t, v, w);
        if constexpr (!LAST) {
            typedef fft_passes<T, LOG2M, INV, PASS + 1, LOG2S + LOG2R, LOG2E> next;
            cpx<T> wn[next::NW];
            if constexpr (!next::LAST) F::template load_twiddles<next::LOG2R, LOG2S + LOG2R>(t, wn, tw, tw_shift); // one pass ahead
            F::template store_pass<LOG2R, LOG2S>(t, v, smem);
            __syncthreads();
            F::load_natural(t, v, smem);
            __syncthreads();
            next::pass(t, v, smem, tw, tw_shift, wn);
        }
    }
#endif

    // host emulation: vs[t] are the registers of thread t
    static void run_host(cpx<T> (*vs)[F::E], cpx<T> *smem, const cpx<T> *tw, int tw_shift)
    {
        for (int t = 0; t < F::NT; t++) {
            cpx<T> w[NW];
            if constexpr (!LAST) F::template load_twiddles<LOG2R, LOG2S>(t, w, tw, tw_shift);
            F::template butterflies<LOG2R, LOG2S, LAST>(t, vs[t], w);
        }
        if constexpr (!LAST) {
            for (int t = 0; t < F::NT; t++) F::template store_pass<LOG2R, LOG2S>(t, vs[t], smem);
            for (int t = 0; t < F::NT; t++) F::load_natural(t, vs[t], smem);
            fft_passes<T, LOG2M, INV, PASS + 1, LOG2S + LOG2R, LOG2E>::run_host(vs, smem, tw, tw_shift);
        }
    }
};

} // namespace bfir
