// Compiles the C++ host mirror classes against libbfir_b200.so and drives them the way the plug-in
// drives the reference (foo_dsp_bfir.cpp:279-345). Without a GPU it checks the failure contract.
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../foo-dsp-bfir_b200/host/brutefir.hpp"
#include "../../foo-dsp-bfir_b200/host/fftw_convolver.hpp"

int main(int argc, char **argv)
{
    const bool expect_gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    const int L = 1024, P = 4, C = 2;
    brutefir filter(L, P, 8, C, BFIR_SAMPLE_FORMAT_FLOAT_LE, BFIR_SAMPLE_FORMAT_FLOAT_LE, 44100, false);
    std::vector<double> h0(L * P, 0.0), h1(L * P, 0.0);
    h0[0] = 1.0;          // identity
    h1[5] = 0.5;          // delay by 5, gain 0.5
    void *coeffs[2] = { h0.data(), h1.data() };
    const int rc = filter.set_coeff(coeffs, 2, L * P, P, 1.0);
    if (!expect_gpu) {
        if (rc == -2 && !filter.is_initialized()) { printf("no device: refused as expected (%s)\n", bfir_last_error()); return 0; }
        printf("unexpected success without a device\n");
        return 1;
    }
    if (rc != 0 || !filter.is_initialized()) { printf("set_coeff failed: %s\n", bfir_last_error()); return 1; }
    std::vector<float> in(L * C), out(L * C), prev(L * C, 0.f);
    double worst = 0;
    for (int b = 0; b < 6; b++) {
        for (int n = 0; n < L * C; n++) in[n] = (float)std::sin(0.01 * (n + 7 * b)) * 0.5f;
        if (filter.run(in.data(), out.data()) != 0) { printf("run failed\n"); return 1; }
        for (int n = 0; n < L; n++) {
            worst = std::fmax(worst, std::fabs(out[n * C] - in[n * C]));
            const float d = n >= 5 ? in[(n - 5) * C + 1] : prev[(L + n - 5) * C + 1];
            worst = std::fmax(worst, std::fabs(out[n * C + 1] - 0.5f * d));
        }
        prev = in;
    }
    filter.check_overflows();
    fftw_convolver conv(L, 4);
    if (!conv.ok() || conv.convolver_cbufsize() != 2 * L * 4) { printf("convolver failed\n"); return 1; }
    printf("host mirror ok, worst abs error %.3g\n", worst);
    return worst < 1e-5 ? 0 : 1;
}
