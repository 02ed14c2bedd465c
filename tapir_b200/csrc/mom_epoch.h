// mom_epoch.h -- host-side plan of the moment-functional epoch integrals (no CUDA in this file).
//
// profile_mom_kernel represents a slowly decaying site (rate < 0.01953125) of a time tile by 32 moments per 128-point
// block,  M[k][hb] = sum_s a_s e^{-c_s T0} (-64 c_s)^k e^{-c_s (128 hb + 64)}  (kernels.cuh).  For such a site the
// first GK21 panel over an epoch [A, B] (centre C, half-length h; tapir/compute.py:50-52 = scipy quad = QUADPACK
// dqk21) is, with y = c h and the rule's pair sums written as their power series (gk21.cuh, c_ak / c_bk),
//
//     V_s = a h e^{-cC} [ C sum_n ak_n y^{2n}  -  h y sum_n bk_n y^{2n} ],
//
// and with  e^{-cC} = e^{-c T0} e^{-c(128 hb + 64)} e^{-c d},  d = C - (T0 + 128 hb + 64) in [-64, 64)  and
// z = -64 c  (so y = -z h/64) it is a polynomial in z times the site's moment weights:
//
//     sum_s V_s = sum_j p_j M[j][hb],   P(z) = h [sum_k (d/64)^k z^k / k!] [sum_n C ak_n (h/64)^{2n} z^{2n}
//                                                                         + h bk_n (h/64)^{2n+1} z^{2n+1}]  mod z^32.
//
// So the epoch sums of compute.py:81-94 over all such sites are one 32-term dot product per (locus, epoch) with
// the moments the profile needs anyway -- no per-site exponential, no second pass over the rates.
//
// What QUADPACK returns for the site must still be this panel with the rounding floor as its error estimate
// (abserr = 50 eps resabs, accepted by dqagse as is).  quad_kernel proves that per site from a lower bound of
// resasc (kernels.cuh, gk21_collapsed); here the same proof is carried out once per EPOCH as a statement about
// rates:  with e0 <= 32 eps resk (y <= 1: Kronrod and Gauss agree to < 1e-19, what dqk21 sees is rounding noise)
// the floor wins and resasc > floor whenever  resasc / resk >= 2.33e-8, and
//
//     resasc/resk >= sum_j wgk_j (|f1_j - m| + |f2_j - m|) / resk >= (2 h / resk_rel) sum_j wgk_j x_j |phi_j(c)|,
//     phi_j(c) = c C sinhc(c h x_j) - cosh(c h x_j),     resk_rel = C Ak(y) - h y Bk(y) <= 3.1 C   (y <= 1),
//
// (f1_j - f2_j = 2 h x_j phi_j(c); all ten node pairs).  Every phi_j vanishes near c = 1/C -- the site whose PI curve
// peaks at the epoch centre -- but at ten different rates, c_j ~ (1 + (h x_j / C)^2 / 3) / C, so the sum stays above the
// threshold unless the epoch is very narrow against its distance from 0 (h/C below ~0.02); when it does dip, the plan
// returns the window of rates around the dip (threshold taken with a 10x margin).  Sites inside the window, above
// r_max or below r_min take the per-site path of quad_kernel for this epoch, and their polynomial value is subtracted
// there.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace tapir {

// The moment class of profile_mom_kernel<RH, KM>: KM Taylor moments per 128-point block represent a site with
// 64 c = 256 r <= x(KM) to ~1e-11 (remainder x^KM / KM! against e^{-x}):  KM = 32: x = 5 (r < 0.01953125),
// KM = 16: x = 1.25 (r < 0.0048828125).  Both bounds are exact binary fractions, so the class test is one comparison of
// the rate's high word.
constexpr int kMomTerms = 32;                 // storage of a plan's coefficients (the largest KM)
inline double mom_class_x(int km) { return km <= 16 ? 1.25 : 5.0; }
inline double mom_rate_bound(int km) { return mom_class_x(km) / 256.0; }
inline uint32_t mom_rate_bound_hi(int km) { return km <= 16 ? 0x3F740000u : 0x3F940000u; }
constexpr double kMomRateBound = 0.01953125;  // KM = 32

struct MomEpochPlan {
    int32_t eligible;    // 0: every site takes the per-site path for this epoch
    int32_t tile, hb;    // anchor block: the 128-point block that contains the epoch centre
    double tb;           // its centre, T0 + 128 hb + 64 (absolute time)
    double p[kMomTerms]; // integral = sum_j p[j] M[j][hb]
    double r_min, r_max; // rates of the moment class that may use the functional: r_min <= r <= r_max ...
    double w_lo, w_hi;   // ... except w_lo <= r <= w_hi (empty when w_lo > w_hi)
};

namespace momdetail {
static const long double kX[10] = {
    0.995657163025808080735527280689003L, 0.973906528517171720077964012084452L, 0.930157491355708226001207180059508L,
    0.865063366688984510732096688423493L, 0.780817726586416897063717578345042L, 0.679409568299024406234327365114874L,
    0.562757134668604683339000099272694L, 0.433395394129247190799265943165784L, 0.294392862701460198131126603103866L,
    0.148874338981631210884826001129720L};
static const long double kW[11] = {
    0.011694638867371874278064396062192L, 0.032558162307964727478818972459390L, 0.054755896574351996031381300244580L,
    0.075039674810919952767043140916190L, 0.093125454583697605535065465083366L, 0.109387158802297641899210590325805L,
    0.123491976262065851077958109585166L, 0.134709217311473325928054001771707L, 0.142775938577060080797094273138717L,
    0.147739104901338491374841515972068L, 0.149445554002916905664936468389821L};

// S(c) = sum_j wgk_j x_j |phi_j(c)|: (resasc / resk) >= 2 h S(c) / (3.1 C)
inline long double phi(int j, long double c, long double C, long double h)
{
    const long double u = c * h * kX[j];
    const long double sinhc = (u < 1e-4L) ? 1.0L + u * u / 6.0L : sinhl(u) / u;
    return c * C * sinhc - coshl(u);
}
inline long double S(long double c, long double C, long double h)
{
    long double s = 0.0L;
    for (int j = 0; j < 10; ++j) s += kW[j] * kX[j] * fabsl(phi(j, c, C, h));
    return s;
}
}  // namespace momdetail

// FP32 mode (profile_f32_kernel): 16 moments with the class 64 c <= 2.5 (rate < 0.009765625) -- remainder 1.3e-6, inside
// that mode's 1e-5 gate.
constexpr double kMomClassF32 = 2.5;
inline uint32_t mom_rate_bound_hi_x(double x) { return x == 2.5 ? 0x3F840000u : (x == 1.25 ? 0x3F740000u : 0x3F940000u); }

// Plan of the epoch [a, b] on a grid of n_tiles tiles of TT points (TT a multiple of 128) for the class of km moments
// with 64 c <= class_x (0: the FP64 class of km, mom_class_x).
inline void mom_epoch_plan(double a, double b, int64_t n_tiles, int TT, int km, MomEpochPlan* out, double class_x = 0.0)
{
    using namespace momdetail;
    std::memset(out, 0, sizeof *out);
    out->w_lo = 1.0;
    out->w_hi = 0.0;
    if (!(a >= 0.0) || !(b > a) || !std::isfinite(b)) return;
    const long double C = 0.5L * ((long double)a + (long double)b), h = 0.5L * ((long double)b - (long double)a);
    // the rounded centre / half-length are what dqk21 (and quad_kernel) use: keep to epochs where they are exact
    if (double(C) != 0.5 * (a + b) || double(h) != 0.5 * (b - a) || (long double)double(C) != C || (long double)double(h) != h) return;
    if (h < 0.5L || h > 32.0L) return;
    if (!(C < (long double)n_tiles * TT)) return;
    const int64_t gb = int64_t(floorl(C / 128.0L));
    const int nb = TT / 128;
    const long double tb = 128.0L * gb + 64.0L, d = C - tb;
    // ---- rate range ----
    const long double xk = class_x > 0.0 ? class_x : mom_class_x(km);
    long double r_max = xk / 256.0L;                          // class bound (exclusive; applied by the class test)
    r_max = fminl(r_max, 1.0L / (4.0L * h));                  // y = c h <= 1
    r_max = fminl(r_max, xk / (4.0L * (fabsl(d) + h)));       // series remainder as for the profile: (|d| + h) c <= x(km)
    r_max = fminl(r_max, 150.0L / (long double)b);            // c b <= 600: every exponential stays a normal number
    const long double r_min = 2.5e-16L;                       // a = 16 r^2 >= 1e-30: result > uflow / (50 eps)
    if (!(r_max > 1e-12L)) return;
    // ---- window of rates where the floor proof may fail ----
    // S(c) >= tau  =>  resasc / resk >= 2.5e-7, more than 10 x the 2.33e-8 the proof needs
    const long double tau = 2.5e-7L * 3.1L * C / (2.0L * h);
    const long double c_max = 4.0L * r_max;
    double w_lo = 1.0, w_hi = 0.0;
    {
        // coarse look at the whole range (S is smooth away from the roots of the phi_j, which all lie in a span of
        // relative width (h/C)^2 / 3 around 1/C)
        const int N = 2048;
        long double s_min = S(0.0L, C, h);
        for (int i = 1; i <= N; ++i) s_min = fminl(s_min, S(c_max * i / N, C, h));
        // roots of the phi_j inside the range: phi_j(0+) = -1; beyond y = 1 they are out of range
        long double root[10];
        int n_root = 0;
        for (int j = 0; j < 10; ++j) {
            if (phi(j, c_max, C, h) <= 0.0L) continue;
            long double lo = 0.0L, hi = c_max;
            for (int it = 0; it < 100; ++it) {
                const long double mid = 0.5L * (lo + hi);
                if (phi(j, mid, C, h) <= 0.0L) lo = mid; else hi = mid;
            }
            root[n_root++] = 0.5L * (lo + hi);
        }
        // between its breakpoints S is a sum of smooth, nearly linear terms: the dip is at a root
        long double c_dip = -1.0L;
        for (int k = 0; k < n_root; ++k) {
            const long double v = S(root[k], C, h);
            if (v < s_min) s_min = v;
            if (c_dip < 0.0L || v < S(c_dip, C, h)) c_dip = root[k];
        }
        if (n_root == 0 && s_min < 4.0L * tau) return;   // no sign change, yet close to the threshold: leave the epoch per site
        if (s_min < 2.0L * tau) {
            if (n_root == 0) return;
            // sub-level set {S < 2 tau} around the dip, found outwards from it (twice the threshold: the set is taken
            // generously, and what lies outside is verified below)
            long double lo = 0.0L, hi = c_dip;
            for (int it = 0; it < 100; ++it) {
                const long double mid = 0.5L * (lo + hi);
                if (S(mid, C, h) >= 2.0L * tau) lo = mid; else hi = mid;
            }
            const long double c_lo = lo;
            lo = c_dip; hi = c_max;
            if (S(c_max, C, h) < 2.0L * tau) lo = c_max;
            else
                for (int it = 0; it < 100; ++it) {
                    const long double mid = 0.5L * (lo + hi);
                    if (S(mid, C, h) >= 2.0L * tau) hi = mid; else lo = mid;
                }
            const long double c_hi = (lo == c_max) ? c_max : hi;
            // verify: outside [c_lo, c_hi] S stays above the threshold proper
            for (int i = 1; i <= 512; ++i) {
                const long double below = c_lo * (1.0L - 0.25L * i / 512.0L), above = c_hi + (c_max - c_hi) * i / 512.0L;
                const long double near_above = c_hi * (1.0L + 0.25L * i / 512.0L);
                if (S(below, C, h) < tau || S(above, C, h) < tau || (near_above < c_max && S(near_above, C, h) < tau)) return;
            }
            w_lo = double(c_lo / 4.0L * (1.0L - 1e-9L));
            w_hi = double(c_hi / 4.0L * (1.0L + 1e-9L));
            if (w_lo <= double(r_min) && w_hi >= double(r_max)) return;   // nothing left
        }
    }
    // ---- the functional ----
    long double E[kMomTerms], G[kMomTerms], fact = 1.0L, dk = 1.0L;
    for (int k = 0; k < kMomTerms; ++k) {
        if (k) { fact *= k; dk *= d / 64.0L; }
        E[k] = dk / fact;
        G[k] = 0.0L;
    }
    const long double r = h / 64.0L;
    long double f2n = 1.0L, rp = 1.0L;  // (2n)!, r^{2n}
    for (int n = 0; 2 * n < kMomTerms; ++n) {
        if (n) { f2n *= (2 * n - 1) * (2 * n); rp *= r * r; }
        long double sa = (n == 0) ? kW[10] : 0.0L, sb = 0.0L;
        for (int j = 0; j < 10; ++j) {
            const long double x2n = powl(kX[j], 2 * n);
            sa += 2.0L * kW[j] * x2n;
            sb += 2.0L * kW[j] * x2n * kX[j] * kX[j];
        }
        G[2 * n] = C * (sa / f2n) * rp;                                              // C ak_n (h/64)^{2n}
        if (2 * n + 1 < kMomTerms) G[2 * n + 1] = h * (sb / (f2n * (2 * n + 1))) * rp * r;   // h bk_n (h/64)^{2n+1}
    }
    for (int j = 0; j < kMomTerms; ++j) {
        long double s = 0.0L;
        for (int k = 0; k <= j; ++k) s += E[k] * G[j - k];
        out->p[j] = (j < km) ? double(h * s) : 0.0;          // P(z) mod z^km
    }
    out->eligible = 1;
    out->tile = int32_t(gb / nb);
    out->hb = int32_t(gb % nb);
    out->tb = double(tb);
    out->r_min = double(r_min);
    out->r_max = double(r_max);
    out->w_lo = w_lo;
    out->w_hi = w_hi;
}

}  // namespace tapir
