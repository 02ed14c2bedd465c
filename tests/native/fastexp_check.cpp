// Host-side accuracy sweep of tapir_b200/csrc/fastexp.cuh against long double expl.
// Prints the worst error in ulp for each entry point; tests/test_fastexp.py asserts on it.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include "../../tapir_b200/csrc/fastexp.cuh"

static double ulp_err(double got, long double want)
{
    if (want == 0.0L) return got == 0.0 ? 0.0 : 1e9;
    if (want > (long double)DBL_MAX) return std::isinf(got) ? 0.0 : 1e9;
    int e;
    std::frexp((double)want, &e);
    const long double ulp = std::ldexp(1.0L, (e - 1 < -1022 ? -1022 : e - 1) - 52);
    return (double)std::fabs(((long double)got - want) / ulp);
}

int main()
{
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    double wb = 0, wpm = 0, wpp = 0, wf = 0;
    const int N = 2000000;
    for (int i = 0; i < N; ++i) {
        // log-uniform magnitudes in [1e-12, 700] interleaved with uniform coverage of [0, 700]
        const double mag = (i & 1) ? 700.0 * u(rng) : std::exp(std::log(1e-12) + u(rng) * (std::log(700.0) - std::log(1e-12)));
        const double x = -mag;
        wb = std::fmax(wb, ulp_err(tapir::exp_bounded(x), expl((long double)x)));
        wb = std::fmax(wb, ulp_err(tapir::exp_bounded(-x), expl((long double)-x)));
        double em, ep;
        tapir::exp_pm_bounded(mag, em, ep);
        wpm = std::fmax(wpm, ulp_err(em, expl((long double)-mag)));
        wpp = std::fmax(wpp, ulp_err(ep, expl((long double)mag)));
        const double xf = (i % 3 == 0) ? -(700.0 + 60.0 * u(rng)) : ((i % 3 == 1) ? 700.0 + 20.0 * u(rng) : x);
        wf = std::fmax(wf, ulp_err(tapir::exp_full(xf), expl((long double)xf)));
    }
    const double specials[] = {0.0, -0.0, -745.2, -746.0, -800.0, -1e300, 709.7, 709.8, 710.0, 1e300,
                               -INFINITY, INFINITY};
    int bad = 0;
    for (double s : specials) {
        const double g = tapir::exp_full(s), w = std::exp(s);
        if (!(g == w || ulp_err(g, expl((long double)s)) <= 1.0)) { ++bad; std::printf("special %g: got %g want %g\n", s, g, w); }
    }
    if (!std::isnan(tapir::exp_full(NAN))) ++bad;
    // exp_pm_small is exp_pm_bounded without the (identity) range reduction: same bits wherever it is used
    int small_diff = 0;
    for (int i = 0; i <= 400000; ++i) {
        const double z = 0.3465 * double(i) / 400000.0;  // z log2(e) < 0.4999
        double a, b, c, d;
        tapir::exp_pm_bounded(z, a, b);
        tapir::exp_pm_small(z, c, d);
        if (std::memcmp(&a, &c, 8) != 0 || std::memcmp(&b, &d, 8) != 0) ++small_diff;
    }
    std::printf("exp_pm_small_diff %d\n", small_diff);
    std::printf("exp_bounded %.3f\nexp_pm_minus %.3f\nexp_pm_plus %.3f\nexp_full %.3f\nspecials_bad %d\n", wb, wpm, wpp, wf, bad);
    return 0;
}
