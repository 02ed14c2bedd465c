// fastexp.cuh -- branch-free double-precision exponentials for the Townsend integrand.
//
// The hot loops only ever need e^x for arguments whose range is known at the call site, so the
// special-case ladders of a general exp() (several predicated branches per call, which show up as
// BSSY/BSYNC/FSEL traffic in the SASS) are replaced by straight-line code:
//
//   exp_bounded(x)        |x| <= 700            one scaling, no checks               17 FP64 ops
//   exp_pm_bounded(z)     0 <= z <= 700         e^-z and e^+z from one reduction     21 FP64 ops for both
//   exp_full(x)           any x (NaN -> NaN)    clamped, two-step scaling, Estrin    ~26 FP64 ops, depth ~12
//
// Method: k = rint(x log2 e); r = x - k ln2 (Cody-Waite, ln2 split so that k*ln2_hi is exact);
// e^r by the degree-13 Taylor polynomial on |r| <= ln2/2 (truncation < 4e-18); result scaled by
// 2^k through the exponent field.  Measured against long double expl: <= 0.87 ulp (<= 2 ulp for the
// shared-reduction pair)
// (tests/test_fastexp.py compiles this header for the host and sweeps it).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define TAPIR_HD __host__ __device__ __forceinline__
#else
#define TAPIR_HD inline
#endif

namespace tapir {

namespace fx {
constexpr double kLog2e = 1.4426950408889634;
constexpr double kLn2Hi = 0x1.62e42fee00000p-1;   // 33 significant bits: k * kLn2Hi is exact for |k| < 2^20
constexpr double kLn2Lo = 0x1.a39ef35793c76p-33;
constexpr double kMagic = 6755399441055744.0;     // 1.5 * 2^52: adding it rounds to nearest integer
// 1/n!  (host: constexpr; device: the same values in constant memory, so that they are constant-bank operands of
// the DFMAs instead of 64-bit immediates that cost two UMOVs per use)
#if defined(__CUDA_ARCH__)
#define TAPIR_FX_C(n) ::tapir::fx::d_coef[n]
#else
#define TAPIR_FX_C(n) ::tapir::fx::h_coef[n]
#endif
constexpr double c2 = 0.5, c3 = 0.16666666666666666, c4 = 0.041666666666666664, c5 = 0.008333333333333333,
                 c6 = 0.001388888888888889, c7 = 0.0001984126984126984, c8 = 2.48015873015873e-05,
                 c9 = 2.7557319223985893e-06, c10 = 2.755731922398589e-07, c11 = 2.505210838544172e-08,
                 c12 = 2.08767569878681e-09, c13 = 1.6059043836821613e-10;

static const double h_coef[14] = {1.0, 1.0, c2, c3, c4, c5, c6, c7, c8, c9, c10, c11, c12, c13};
#if defined(__CUDACC__)
__device__ __constant__ double d_coef[14] = {1.0, 1.0, c2, c3, c4, c5, c6, c7, c8, c9, c10, c11, c12, c13};
#endif

TAPIR_HD double fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return ::fma(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

// 2^k for -1022 <= k <= 1023
TAPIR_HD double pow2i(int k)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double((k + 1023) << 20, 0);
#else
    const uint64_t bits = uint64_t(k + 1023) << 52;
    double d;
    std::memcpy(&d, &bits, sizeof d);
    return d;
#endif
}

// p * 2^k by adding k to p's exponent field: exact while p and the result are normal numbers (the bounded
// variants guarantee it); one integer add instead of building 2^k and an FP64 multiply
TAPIR_HD double scale_pow2(double p, int k)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    uint64_t bits;
    std::memcpy(&bits, &p, sizeof bits);
    bits += uint64_t(int64_t(k)) << 52;
    std::memcpy(&p, &bits, sizeof p);
    return p;
#endif
}

// integer nearest to t (|t| < 2^31) as both double and int
TAPIR_HD void round_int(double t, double& kd, int& ki)
{
    const double m = t + kMagic;
#if defined(__CUDA_ARCH__)
    ki = __double2loint(m);
#else
    uint64_t bits;
    std::memcpy(&bits, &m, sizeof bits);
    ki = int(uint32_t(bits));
#endif
    kd = m - kMagic;
}

TAPIR_HD double reduce(double x, double kd) { return fma_(kd, -kLn2Lo, fma_(kd, -kLn2Hi, x)); }

// e^r on |r| <= 0.35
TAPIR_HD double poly(double r)
{
    double p = TAPIR_FX_C(13);
    p = fma_(p, r, TAPIR_FX_C(12)); p = fma_(p, r, TAPIR_FX_C(11)); p = fma_(p, r, TAPIR_FX_C(10));
    p = fma_(p, r, TAPIR_FX_C(9));  p = fma_(p, r, TAPIR_FX_C(8));  p = fma_(p, r, TAPIR_FX_C(7));
    p = fma_(p, r, TAPIR_FX_C(6));  p = fma_(p, r, TAPIR_FX_C(5));  p = fma_(p, r, TAPIR_FX_C(4));
    p = fma_(p, r, TAPIR_FX_C(3));  p = fma_(p, r, TAPIR_FX_C(2));
    p = fma_(p, r, 1.0);
    return fma_(p, r, 1.0);
}

// the same polynomial in Estrin form: 3 more operations, dependent depth 6 instead of 14 (for callers
// that have a single exponential in flight per thread)
TAPIR_HD double poly_estrin(double r)
{
    // e^r = 1 + (r + r^2 P(r)),  P = c2 + c3 r + ... + c13 r^11 evaluated pairwise
    const double r2 = r * r;
    const double p0 = fma_(TAPIR_FX_C(3), r, TAPIR_FX_C(2)), p1 = fma_(TAPIR_FX_C(5), r, TAPIR_FX_C(4));
    const double p2 = fma_(TAPIR_FX_C(7), r, TAPIR_FX_C(6)), p3 = fma_(TAPIR_FX_C(9), r, TAPIR_FX_C(8));
    const double p4 = fma_(TAPIR_FX_C(11), r, TAPIR_FX_C(10)), p5 = fma_(TAPIR_FX_C(13), r, TAPIR_FX_C(12));
    const double r4 = r2 * r2;
    const double q0 = fma_(p1, r2, p0), q1 = fma_(p3, r2, p2), q2 = fma_(p5, r2, p4);
    const double r8 = r4 * r4;
    const double P = fma_(q2, r8, fma_(q1, r4, q0));
    return 1.0 + fma_(r2, P, r);
}
}  // namespace fx

// e^x for |x| <= 700 (no range checks; the result is a normal number)
TAPIR_HD double exp_bounded(double x)
{
    double kd;
    int ki;
    fx::round_int(x * fx::kLog2e, kd, ki);
    return fx::scale_pow2(fx::poly(fx::reduce(x, kd)), ki);
}

// exp_bounded with the polynomial in Estrin form: dependent depth 12 instead of 19, for callers with one exponential
// in flight per thread whose FP64 operations queue behind other warps' DMMAs (profile_mom_kernel's table build)
TAPIR_HD double exp_bounded_short(double x)
{
    double kd;
    int ki;
    fx::round_int(x * fx::kLog2e, kd, ki);
    return fx::scale_pow2(fx::poly_estrin(fx::reduce(x, kd)), ki);
}

// N independent exp_bounded evaluations advanced in lock step, so that the N dependent Horner chains are
// interleaved in the instruction stream (the compiler otherwise emits them one after the other and every
// DFMA waits out the 8-cycle latency of the previous one)
template <int N>
TAPIR_HD void exp_bounded_n(const double (&x)[N], double (&y)[N])
{
    double r[N], p[N];
    int ki[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double kd;
        fx::round_int(x[i] * fx::kLog2e, kd, ki[i]);
        r[i] = fx::reduce(x[i], kd);
        p[i] = TAPIR_FX_C(13);
    }
#pragma unroll
    for (int k = 12; k >= 2; --k)
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = fx::fma_(p[i], r[i], TAPIR_FX_C(k));
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fx::fma_(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = fx::scale_pow2(fx::fma_(p[i], r[i], 1.0), ki[i]);
}

// (e^-z, e^+z) for 0 <= z <= 700 from one range reduction: e^{+/-r} = E(r^2) +/- r O(r^2)
TAPIR_HD void exp_pm_bounded(double z, double& em, double& ep)
{
    using namespace fx;
    double kd;
    int ki;
    round_int(z * kLog2e, kd, ki);
    const double r = reduce(z, kd);
    const double s = r * r;
    double ev = TAPIR_FX_C(12), od = TAPIR_FX_C(13);
    ev = fma_(ev, s, TAPIR_FX_C(10)); od = fma_(od, s, TAPIR_FX_C(11));
    ev = fma_(ev, s, TAPIR_FX_C(8));  od = fma_(od, s, TAPIR_FX_C(9));
    ev = fma_(ev, s, TAPIR_FX_C(6));  od = fma_(od, s, TAPIR_FX_C(7));
    ev = fma_(ev, s, TAPIR_FX_C(4));  od = fma_(od, s, TAPIR_FX_C(5));
    ev = fma_(ev, s, TAPIR_FX_C(2));  od = fma_(od, s, TAPIR_FX_C(3));
    ev = fma_(ev, s, 1.0); od = fma_(od, s, 1.0);
    ep = scale_pow2(fma_(od, r, ev), ki);
    em = scale_pow2(fma_(od, -r, ev), -ki);
}

// The same pair for 0 <= z with z log2(e) < 1/2: the reduction above finds k = 0 and r = z exactly, so it is left
// out; the results are bit-identical to exp_pm_bounded's (17 FP64 operations for both).
TAPIR_HD void exp_pm_small(double z, double& em, double& ep)
{
    using namespace fx;
    const double s = z * z;
    double ev = TAPIR_FX_C(12), od = TAPIR_FX_C(13);
    ev = fma_(ev, s, TAPIR_FX_C(10)); od = fma_(od, s, TAPIR_FX_C(11));
    ev = fma_(ev, s, TAPIR_FX_C(8));  od = fma_(od, s, TAPIR_FX_C(9));
    ev = fma_(ev, s, TAPIR_FX_C(6));  od = fma_(od, s, TAPIR_FX_C(7));
    ev = fma_(ev, s, TAPIR_FX_C(4));  od = fma_(od, s, TAPIR_FX_C(5));
    ev = fma_(ev, s, TAPIR_FX_C(2));  od = fma_(od, s, TAPIR_FX_C(3));
    ev = fma_(ev, s, 1.0); od = fma_(od, s, 1.0);
    ep = fma_(od, z, ev);
    em = fma_(od, -z, ev);
}

// e^x for any x: +inf above the overflow threshold, gradual underflow to 0 below, NaN -> NaN
TAPIR_HD double exp_full(double x)
{
    const double xc = fmin(fmax(x, -760.0), 760.0);
    double kd;
    int ki;
    fx::round_int(xc * fx::kLog2e, kd, ki);
    const double p = fx::poly_estrin(fx::reduce(xc, kd));
    const int k1 = ki >> 1;
    const double v = (p * fx::pow2i(k1)) * fx::pow2i(ki - k1);
    return (x != x) ? x : v;
}

}  // namespace tapir
