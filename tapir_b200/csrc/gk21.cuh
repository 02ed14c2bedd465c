// gk21.cuh -- device-side QUADPACK pieces for the Townsend integrand.
//
// Replaces scipy.integrate.quad as called by tapir/compute.py:50-52
// (get_integral_over_times): QAGS (dqagse) with the 21-point Gauss-Kronrod rule (dqk21),
// epsabs = epsrel = 1.49e-8, limit = 50.  The node/weight tables and the order in which the
// Gauss and Kronrod sums are accumulated follow QUADPACK (Piessens et al. 1983) so that
// `result` and `abserr` land on the reference's values to rounding.
#pragma once
#include <cfloat>
#include <cstdint>

namespace tapir {

constexpr double kQuadEps = 1.49e-8;  // scipy quad defaults (compute.py:52 passes none)
constexpr int kQuadLimit = 50;
constexpr double kEpmach = DBL_EPSILON;
constexpr double kUflow = DBL_MIN;
constexpr double kOflow = DBL_MAX;

// Kronrod abscissae x_0..x_9 (x_10 = 0 is the centre), Kronrod weights w_0..w_10, and the Gauss
// weight that goes with every odd-indexed abscissa (0 for the Kronrod-only nodes).
__device__ __constant__ double c_xgk[10] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720};
__device__ __constant__ double c_wgk[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};
__device__ __constant__ double c_wg[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769};

// Products of the rule's constants for the collapsed panel of quad_kernel (kernels.cuh, gk21_collapsed):
// c_wx[j] = wgk[j] x[j]; c_dl[j] = wgk[j] - wg[(j-1)/2] on the Gauss nodes (odd j), wgk[j] elsewhere, so that
// resk - resg = wgk[10] f(centre) + sum_j c_dl[j] (f1_j + f2_j); c_dlx[j] = c_dl[j] x[j].  Derived in 60-digit
// decimal arithmetic from the doubles above and rounded once.
__device__ __constant__ double c_wx[10] = {0.01164385095729883,  0.03170860682824855, 0.05093160739453187, 0.06491407372718101,
                                           0.07271400573536937,  0.07431868233932587, 0.06949599071580352, 0.0583823543295484,
                                           0.04203221728258857,  0.02199456158392464};
__device__ __constant__ double c_dl[10] = {0.011694638867371874, -0.03411318200072341, 0.054755896574351995, -0.07441167433966063,
                                           0.0931254545836976,   -0.1096992037136844,  0.12349197626206584,  -0.134557501998523,
                                           0.14277593857706009,  -0.14778511981341438};
__device__ __constant__ double c_dlx[10] = {0.01164385095729883,  -0.033223050658999005, 0.05093160739453187, -0.06437081352523115,
                                            0.07271400573536937,  -0.07453068863786105,  0.06949599071580352, -0.05831660161169685,
                                            0.04203221728258857,  -0.022001412023543238};

// Power series of the Kronrod pair sums in y = c h (quad_kernel's slow-site branch, y x_0 log2(e) < 1/2):
//   wgk_10 + sum_j wgk_j (e^{y x_j} + e^{-y x_j})  = sum_n c_ak[n] y^{2n},     c_ak[n] = (2 sum_j wgk_j x_j^{2n} [+ wgk_10]) / (2n)!
//   sum_j wgk_j x_j (e^{y x_j} - e^{-y x_j})       = sum_n c_bk[n] y^{2n+1},   c_bk[n] = 2 sum_j wgk_j x_j^{2n+2} / (2n+1)!
// (the rule's moments, 80-digit decimal arithmetic on the doubles above; they agree with the moments of dx on [-1,1],
// 2/(2n+1), to the last bit or two -- the rule is exact far beyond degree 14).  Seven terms: truncation < 3e-19 at
// the largest y of the branch.
__device__ __constant__ double c_ak[7] = {2.0, 0.33333333333333337, 0.016666666666666666, 0.0003968253968253969,
                                          5.511463844797179e-06, 5.010421677088345e-08, 3.2118087673643237e-10};
__device__ __constant__ double c_bk[7] = {0.6666666666666667, 0.06666666666666667, 0.002380952380952381, 4.4091710758377435e-05,
                                          5.010421677088345e-07, 3.854170520837188e-09, 2.1412058449095492e-11};

// Host copies of the abscissae (the host precomputes centre -/+ h*x_j per epoch).
static const double h_xgk[10] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720};

// compute.py:48 with the per-site invariants hoisted: a = 16 r^2, c = 4 r.
__device__ __forceinline__ double townsend_ac(double t, double a, double c) { return a * t * exp(-(c * t)); }

// dqk21's error heuristic from (resk - resg), resabs and resasc.
__device__ __forceinline__ double gk_error(double resk, double resg, double hlgth, double resabs, double resasc)
{
    double err = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && err != 0.0) {
        double n = 200.0 * err;
        if (n >= resasc) {
            err = resasc;  // min(1, x^1.5) == 1
        } else {
            double x = n / resasc;
            err = resasc * (x * sqrt(x));
        }
    }
    if (resabs > kUflow / (50.0 * kEpmach)) err = fmax((kEpmach * 50.0) * resabs, err);
    return err;
}

// One GK21 panel on [a,b], integrand evaluated directly (21 exp calls).  Accumulation order as
// dqk21: centre, Gauss nodes (odd Kronrod indices), then the Kronrod-only nodes.
__device__ __noinline__ void gk21_direct(double a, double b, double ca, double cc, double* result, double* abserr,
                                         double* resabs_out, double* resasc_out)
{
    double fv1[10], fv2[10];
    const double centr = 0.5 * (a + b);
    const double hlgth = 0.5 * (b - a);
    const double dhlgth = fabs(hlgth);
    double resg = 0.0;
    const double fc = townsend_ac(centr, ca, cc);
    double resk = c_wgk[10] * fc;
    double resabs = fabs(resk);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jtw = 2 * j + 1;
        const double absc = hlgth * c_xgk[jtw];
        const double f1 = townsend_ac(centr - absc, ca, cc);
        const double f2 = townsend_ac(centr + absc, ca, cc);
        const double fsum = f1 + f2;
        fv1[jtw] = f1; fv2[jtw] = f2;
        resg += c_wg[j] * fsum;
        resk += c_wgk[jtw] * fsum;
        resabs += c_wgk[jtw] * (fabs(f1) + fabs(f2));
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jt = 2 * j;
        const double absc = hlgth * c_xgk[jt];
        const double f1 = townsend_ac(centr - absc, ca, cc);
        const double f2 = townsend_ac(centr + absc, ca, cc);
        const double fsum = f1 + f2;
        fv1[jt] = f1; fv2[jt] = f2;
        resk += c_wgk[jt] * fsum;
        resabs += c_wgk[jt] * (fabs(f1) + fabs(f2));
    }
    const double reskh = resk * 0.5;
    double resasc = c_wgk[10] * fabs(fc - reskh);
#pragma unroll
    for (int j = 0; j < 10; ++j) resasc += c_wgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * hlgth;
    resabs *= dhlgth;
    resasc *= dhlgth;
    *abserr = gk_error(resk, resg, hlgth, resabs, resasc);
    *resabs_out = resabs;
    *resasc_out = resasc;
}

// dqagse's acceptance test after the first panel: true when the single-panel (result, abserr)
// is what QAGS returns.
__device__ __forceinline__ bool qags_first_panel_final(double result, double abserr, double resabs, double resasc)
{
    const double errbnd = fmax(kQuadEps, kQuadEps * fabs(result));
    const bool roundoff = (abserr <= 100.0 * kEpmach * resabs) && (abserr > errbnd);  // ier = 2
    return roundoff || (abserr <= errbnd && abserr != resasc) || abserr == 0.0;
}

// dqpsrt (1-based indices kept so the logic reads like QUADPACK's).
__device__ inline void qpsrt(int limit, int last, int& maxerr, double& ermax, const double* elist, int* iord,
                             int& nrmax)
{
    if (last <= 2) {
        iord[1] = 1; iord[2] = 2;
    } else {
        const double errmax = elist[maxerr];
        if (nrmax != 1) {
            const int ido = nrmax - 1;
            for (int i = 1; i <= ido; ++i) {
                const int isucc = iord[nrmax - 1];
                if (errmax <= elist[isucc]) break;
                iord[nrmax] = isucc;
                --nrmax;
            }
        }
        int jupbn = last;
        if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
        const double errmin = elist[last];
        const int jbnd = jupbn - 1;
        const int ibeg = nrmax + 1;
        bool placed = false;
        for (int i = ibeg; i <= jbnd; ++i) {
            int isucc = iord[i];
            if (errmax >= elist[isucc]) {
                iord[i - 1] = maxerr;
                int k = jbnd;
                bool low = false;
                for (int j = i; j <= jbnd; ++j) {
                    isucc = iord[k];
                    if (errmin < elist[isucc]) { iord[k + 1] = last; low = true; break; }
                    iord[k + 1] = isucc;
                    --k;
                }
                if (!low) iord[i] = last;
                placed = true;
                break;
            }
            iord[i - 1] = isucc;
        }
        if (!placed) {
            iord[jbnd] = maxerr;
            iord[jupbn] = last;
        }
    }
    maxerr = iord[nrmax];
    ermax = elist[maxerr];
}

// dqelg: epsilon algorithm, table epstab[1..n].
__device__ inline void qelg(int& n, double* epstab, double& result, double& abserr, double* res3la, int& nres)
{
    const int limexp = 50;
    ++nres;
    abserr = kOflow;
    result = epstab[n];
    if (n >= 3) {
        epstab[n + 2] = epstab[n];
        const int newelm = (n - 1) / 2;
        epstab[n] = kOflow;
        const int num = n;
        int k1 = n;
        bool converged = false;
        for (int i = 1; i <= newelm; ++i) {
            const int k2 = k1 - 1, k3 = k1 - 2;
            double res = epstab[k1 + 2];
            const double e0 = epstab[k3], e1 = epstab[k2], e2 = res;
            const double e1abs = fabs(e1);
            const double delta2 = e2 - e1, err2 = fabs(delta2);
            const double tol2 = fmax(fabs(e2), e1abs) * kEpmach;
            const double delta3 = e1 - e0, err3 = fabs(delta3);
            const double tol3 = fmax(e1abs, fabs(e0)) * kEpmach;
            if (err2 <= tol2 && err3 <= tol3) {
                result = res;
                abserr = err2 + err3;
                converged = true;
                break;
            }
            const double e3 = epstab[k1];
            epstab[k1] = e1;
            const double delta1 = e1 - e3, err1 = fabs(delta1);
            const double tol1 = fmax(e1abs, fabs(e3)) * kEpmach;
            if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) { n = i + i - 1; break; }
            const double ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
            const double epsinf = fabs(ss * e1);
            if (epsinf <= 1.0e-4) { n = i + i - 1; break; }
            res = e1 + 1.0 / ss;
            epstab[k1] = res;
            k1 -= 2;
            const double error = err2 + fabs(res - e2) + err3;
            if (error <= abserr) { abserr = error; result = res; }
        }
        if (!converged) {
            if (n == limexp) n = 2 * (limexp / 2) - 1;
            int ib = ((num / 2) * 2 == num) ? 2 : 1;
            const int ie = newelm + 1;
            for (int i = 1; i <= ie; ++i) { epstab[ib] = epstab[ib + 2]; ib += 2; }
            if (num != n) {
                int indx = num - n + 1;
                for (int i = 1; i <= n; ++i) { epstab[i] = epstab[indx]; ++indx; }
            }
            if (nres < 4) {
                res3la[nres] = result;
                abserr = kOflow;
            } else {
                abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
                res3la[1] = res3la[2];
                res3la[2] = res3la[3];
                res3la[3] = result;
            }
        }
    }
    abserr = fmax(abserr, 5.0 * kEpmach * fabs(result));
}

// Full QAGS on [a,b] for one site; `last_out` = number of subintervals used.  Scratch lives in
// thread-local memory; only called for the (rare) sites whose first panel is not accepted.
__device__ __noinline__ void qags_device(double a, double b, double ca, double cc, double* result_out,
                                         double* abserr_out, int* last_out)
{
    const int limit = kQuadLimit;
    double alist[kQuadLimit + 1], blist[kQuadLimit + 1], rlist[kQuadLimit + 1], elist[kQuadLimit + 1];
    int iord[kQuadLimit + 2];
    double rlist2[53], res3la[4];
    double result, abserr, defabs, resabs;
    gk21_direct(a, b, ca, cc, &result, &abserr, &defabs, &resabs);
    const double dres = fabs(result);
    double errbnd = fmax(kQuadEps, kQuadEps * dres);
    int last = 1, ier = 0, ierro = 0;
    alist[1] = a; blist[1] = b; rlist[1] = result; elist[1] = abserr; iord[1] = 1;
    if (abserr <= 100.0 * kEpmach * defabs && abserr > errbnd) ier = 2;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) {
        *result_out = result; *abserr_out = abserr; *last_out = 1;
        return;
    }
    rlist2[1] = result;
    double errmax = abserr, area = result, errsum = abserr;
    abserr = kOflow;
    int maxerr = 1, nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0;
    bool extrap = false, noext = false;
    int iroff1 = 0, iroff2 = 0, iroff3 = 0;
    const int ksgn = (dres >= (1.0 - 50.0 * kEpmach) * defabs) ? 1 : -1;
    double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0;
    bool sum_intervals = false;

    for (last = 2; last <= limit; ++last) {
        const double a1 = alist[maxerr];
        const double b1 = 0.5 * (alist[maxerr] + blist[maxerr]);
        const double a2 = b1, b2 = blist[maxerr];
        const double erlast = errmax;
        double area1, area2, error1, error2, defab1, defab2, tmp;
        gk21_direct(a1, b1, ca, cc, &area1, &error1, &tmp, &defab1);
        gk21_direct(a2, b2, ca, cc, &area2, &error2, &tmp, &defab2);
        const double area12 = area1 + area2;
        const double erro12 = error1 + error2;
        errsum = errsum + erro12 - errmax;
        area = area + area12 - rlist[maxerr];
        if (defab1 != error1 && defab2 != error2) {
            if (fabs(rlist[maxerr] - area12) <= 1.0e-5 * fabs(area12) && erro12 >= 0.99 * errmax) {
                if (extrap) ++iroff2; else ++iroff1;
            }
            if (last > 10 && erro12 > errmax) ++iroff3;
        }
        rlist[maxerr] = area1;
        rlist[last] = area2;
        errbnd = fmax(kQuadEps, kQuadEps * fabs(area));
        if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
        if (iroff2 >= 5) ierro = 3;
        if (last == limit) ier = 1;
        if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * kEpmach) * (fabs(a2) + 1000.0 * kUflow)) ier = 4;
        if (error2 > error1) {
            alist[maxerr] = a2; alist[last] = a1; blist[last] = b1;
            rlist[maxerr] = area2; rlist[last] = area1;
            elist[maxerr] = error2; elist[last] = error1;
        } else {
            alist[last] = a2; blist[maxerr] = b1; blist[last] = b2;
            elist[maxerr] = error1; elist[last] = error2;
        }
        qpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
        if (errsum <= errbnd) { sum_intervals = true; break; }
        if (ier != 0) break;
        if (last == 2) {
            small = fabs(b - a) * 0.375;
            erlarg = errsum;
            ertest = errbnd;
            rlist2[2] = area;
            continue;
        }
        if (noext) continue;
        erlarg -= erlast;
        if (fabs(b1 - a1) > small) erlarg += erro12;
        if (!extrap) {
            if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
            extrap = true;
            nrmax = 2;
        }
        if (ierro != 3 && erlarg > ertest) {
            bool jumped = false;
            const int id = nrmax;
            int jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            for (int k = id; k <= jupbnd; ++k) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) { jumped = true; break; }
                ++nrmax;
            }
            if (jumped) continue;
        }
        ++numrl2;
        rlist2[numrl2] = area;
        double reseps, abseps;
        qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
        ++ktmin;
        if (ktmin > 5 && abserr < 1.0e-3 * errsum) ier = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = fmax(kQuadEps, kQuadEps * fabs(reseps));
            if (abserr <= ertest) break;
        }
        if (numrl2 == 1) noext = true;
        if (ier == 5) break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = false;
        small *= 0.5;
        erlarg = errsum;
    }
    if (last > limit) last = limit;

    if (!sum_intervals) {
        if (abserr == kOflow) {
            sum_intervals = true;
        } else if (ier + ierro != 0) {
            if (ierro == 3) abserr += correc;
            if (ier == 0) ier = 3;
            if (result != 0.0 && area != 0.0) {
                if (abserr / fabs(result) > errsum / fabs(area)) sum_intervals = true;
            } else if (abserr > errsum) {
                sum_intervals = true;
            }
            // (the divergence test of dqagse only sets ier, which quad() turns into a warning)
        }
    }
    if (sum_intervals) {
        result = 0.0;
        for (int k = 1; k <= last; ++k) result += rlist[k];
        abserr = errsum;
    }
    *result_out = result;
    *abserr_out = abserr;
    *last_out = last;
}

}  // namespace tapir
