/*
 * oracle/quadpack_port.c -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Plain-C, single-threaded CPU restatement of the arithmetic behind tapir's
 * Townsend phylogenetic-informativeness hot path.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * the library built from this file; the product path (tapir_b200/) never does.
 *
 * What is restated and where it lives in the reference:
 *   townsend_f          tapir/compute.py:46-48   get_townsend_pi
 *   oracle_qk21         scipy.integrate.quad -> QUADPACK dqk21 (21-pt Gauss-Kronrod);
 *                       call site tapir/compute.py:50-52 get_integral_over_times
 *   oracle_qags         QUADPACK dqagse (+ dqelg, dqpsrt) with quad()'s defaults
 *                       epsabs = epsrel = 1.49e-8, limit = 50 (compute.py:52)
 *   oracle_profile      compute.py:46-48 + bin/tapir_compute.py:119 (nansum over sites)
 *   oracle_epochs       compute.py:81-94 get_net_integral_for_epochs on
 *                       rates[isfinite(rates)] (bin/tapir_compute.py:122)
 *
 * The integration algorithm is a third-party dependency of the reference
 * (scipy >= 0.9.0, setup.py:26; QUADPACK by Piessens, de Doncker-Kapenga,
 * Ueberhuber, Kahaner, 1983) whose source is NOT under /root/reference.  This
 * file restates the published algorithm; parity is pinned by
 * tests/test_oracle_quadpack.py, which compares (result, abserr, neval) against
 * scipy.integrate.quad itself bit for bit, and by tests/golden/ fixtures
 * generated from the reference's own compute.py.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, so the
 * operation order below is the operation order executed).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stddef.h>

#define QP_LIMIT_MAX 64

/* compute.py:48 -- evaluation order ((16*r^2)*t)*exp(-((4*r)*t)) */
static double townsend_f(double t, double rate)
{
    return 16.0 * (rate * rate) * t * exp(-(4.0 * rate * t));
}

double oracle_townsend_pi(double t, double rate) { return townsend_f(t, rate); }

/* QUADPACK dqk21 node / weight tables (Kronrod abscissae, Kronrod weights, Gauss weights). */
static const double XGK[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
    0.0};
static const double WGK[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821};
static const double WG[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769};

/* One 21-point Gauss-Kronrod panel on [a,b] for f(t) = townsend_f(t, rate). */
void oracle_qk21(double a, double b, double rate, double *result, double *abserr,
                 double *resabs, double *resasc)
{
    const double epmach = DBL_EPSILON, uflow = DBL_MIN;
    double fv1[10], fv2[10];
    double centr = 0.5 * (a + b);
    double hlgth = 0.5 * (b - a);
    double dhlgth = fabs(hlgth);
    double resg = 0.0;
    double fc = townsend_f(centr, rate);
    double resk = WGK[10] * fc;
    double rabs = fabs(resk);
    int j;
    /* Gauss nodes (odd Kronrod indices) first ... */
    for (j = 0; j < 5; ++j) {
        int jtw = 2 * j + 1;
        double absc = hlgth * XGK[jtw];
        double f1 = townsend_f(centr - absc, rate);
        double f2 = townsend_f(centr + absc, rate);
        double fsum = f1 + f2;
        fv1[jtw] = f1; fv2[jtw] = f2;
        resg += WG[j] * fsum;
        resk += WGK[jtw] * fsum;
        rabs += WGK[jtw] * (fabs(f1) + fabs(f2));
    }
    /* ... then the Kronrod-only nodes. */
    for (j = 0; j < 5; ++j) {
        int jtwm1 = 2 * j;
        double absc = hlgth * XGK[jtwm1];
        double f1 = townsend_f(centr - absc, rate);
        double f2 = townsend_f(centr + absc, rate);
        double fsum = f1 + f2;
        fv1[jtwm1] = f1; fv2[jtwm1] = f2;
        resk += WGK[jtwm1] * fsum;
        rabs += WGK[jtwm1] * (fabs(f1) + fabs(f2));
    }
    double reskh = resk * 0.5;
    double rasc = WGK[10] * fabs(fc - reskh);
    for (j = 0; j < 10; ++j)
        rasc += WGK[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * hlgth;
    rabs *= dhlgth;
    rasc *= dhlgth;
    double err = fabs((resk - resg) * hlgth);
    if (rasc != 0.0 && err != 0.0) {
        double r = pow(200.0 * err / rasc, 1.5);
        err = rasc * (r < 1.0 ? r : 1.0);
    }
    if (rabs > uflow / (50.0 * epmach)) {
        double fl = (epmach * 50.0) * rabs;
        if (fl > err) err = fl;
    }
    *abserr = err; *resabs = rabs; *resasc = rasc;
}

/* QUADPACK dqpsrt: keep iord[] sorted by decreasing error estimate (1-based logic kept). */
static void qpsrt(int limit, int last, int *maxerr, double *ermax, const double *elist,
                  int *iord, int *nrmax)
{
    int i, j, k, ido, isucc, jbnd, jupbn, ibeg;
    double errmax, errmin;
    if (last <= 2) {
        iord[1] = 1; iord[2] = 2;
        goto done;
    }
    errmax = elist[*maxerr];
    if (*nrmax != 1) {
        ido = *nrmax - 1;
        for (i = 1; i <= ido; ++i) {
            isucc = iord[*nrmax - 1];
            if (errmax <= elist[isucc]) break;
            iord[*nrmax] = isucc;
            --(*nrmax);
        }
    }
    jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    errmin = elist[last];
    jbnd = jupbn - 1;
    ibeg = *nrmax + 1;
    for (i = ibeg; i <= jbnd; ++i) {
        isucc = iord[i];
        if (errmax >= elist[isucc]) {
            /* insert errmin by traversing the list bottom-up */
            iord[i - 1] = *maxerr;
            k = jbnd;
            for (j = i; j <= jbnd; ++j) {
                isucc = iord[k];
                if (errmin < elist[isucc]) { iord[k + 1] = last; goto done; }
                iord[k + 1] = isucc;
                --k;
            }
            iord[i] = last;
            goto done;
        }
        iord[i - 1] = isucc;
    }
    iord[jbnd] = *maxerr;
    iord[jupbn] = last;
done:
    *maxerr = iord[*nrmax];
    *ermax = elist[*maxerr];
}

/* QUADPACK dqelg: Wynn epsilon algorithm on the table epstab[1..n] (1-based). */
static void qelg(int *n, double *epstab, double *result, double *abserr, double *res3la, int *nres)
{
    const double epmach = DBL_EPSILON, oflow = DBL_MAX;
    const int limexp = 50;
    int i, ib, ib2, ie, indx, k1, k2, k3, newelm, num;
    double delta1, delta2, delta3, epsinf, error, err1, err2, err3, e0, e1, e1abs, e2, e3, res, ss,
        tol1, tol2, tol3;
    ++(*nres);
    *abserr = oflow;
    *result = epstab[*n];
    if (*n < 3) goto fin;
    epstab[*n + 2] = epstab[*n];
    newelm = (*n - 1) / 2;
    epstab[*n] = oflow;
    num = *n;
    k1 = *n;
    for (i = 1; i <= newelm; ++i) {
        k2 = k1 - 1;
        k3 = k1 - 2;
        res = epstab[k1 + 2];
        e0 = epstab[k3];
        e1 = epstab[k2];
        e2 = res;
        e1abs = fabs(e1);
        delta2 = e2 - e1;
        err2 = fabs(delta2);
        tol2 = fmax(fabs(e2), e1abs) * epmach;
        delta3 = e1 - e0;
        err3 = fabs(delta3);
        tol3 = fmax(e1abs, fabs(e0)) * epmach;
        if (err2 <= tol2 && err3 <= tol3) {
            /* e0, e1, e2 equal to machine accuracy: convergence assumed */
            *result = res;
            *abserr = err2 + err3;
            goto fin;
        }
        e3 = epstab[k1];
        epstab[k1] = e1;
        delta1 = e1 - e3;
        err1 = fabs(delta1);
        tol1 = fmax(e1abs, fabs(e3)) * epmach;
        if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
            *n = i + i - 1;
            break;
        }
        ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
        epsinf = fabs(ss * e1);
        if (epsinf <= 1.0e-4) {
            *n = i + i - 1;
            break;
        }
        res = e1 + 1.0 / ss;
        epstab[k1] = res;
        k1 -= 2;
        error = err2 + fabs(res - e2) + err3;
        if (error <= *abserr) {
            *abserr = error;
            *result = res;
        }
    }
    /* shift the table */
    if (*n == limexp) *n = 2 * (limexp / 2) - 1;
    ib = ((num / 2) * 2 == num) ? 2 : 1;
    ie = newelm + 1;
    for (i = 1; i <= ie; ++i) {
        ib2 = ib + 2;
        epstab[ib] = epstab[ib2];
        ib = ib2;
    }
    if (num != *n) {
        indx = num - *n + 1;
        for (i = 1; i <= *n; ++i) {
            epstab[i] = epstab[indx];
            ++indx;
        }
    }
    if (*nres < 4) {
        res3la[*nres] = *result;
        *abserr = oflow;
    } else {
        *abserr = fabs(*result - res3la[3]) + fabs(*result - res3la[2]) + fabs(*result - res3la[1]);
        res3la[1] = res3la[2];
        res3la[2] = res3la[3];
        res3la[3] = *result;
    }
fin:
    *abserr = fmax(*abserr, 5.0 * epmach * fabs(*result));
}

/*
 * QUADPACK dqagse on f(t) = townsend_f(t, rate).  Returns ier (QUADPACK numbering).
 * limit <= QP_LIMIT_MAX.  scipy.integrate.quad(f, a, b) == oracle_qags(a, b, rate,
 * 1.49e-8, 1.49e-8, 50, ...).
 */
int oracle_qags(double a, double b, double rate, double epsabs, double epsrel, int limit,
                double *result_out, double *abserr_out, int *neval_out, int *last_out)
{
    const double epmach = DBL_EPSILON, uflow = DBL_MIN, oflow = DBL_MAX;
    double alist[QP_LIMIT_MAX + 1], blist[QP_LIMIT_MAX + 1], rlist[QP_LIMIT_MAX + 1],
        elist[QP_LIMIT_MAX + 1];
    int iord[QP_LIMIT_MAX + 2];
    double rlist2[53], res3la[4];
    double result = 0.0, abserr = 0.0, defabs, resabs, dres, errbnd;
    double area, area1, area2, area12, a1, a2, b1, b2, correc = 0.0, defab1, defab2, erlarg = 0.0,
        erlast, errmax, error1, error2, erro12, errsum, ertest = 0.0, reseps, abseps, small = 0.0;
    int ier = 0, ierro = 0, last, maxerr, nrmax, nres, numrl2, ktmin, extrap, noext, iroff1, iroff2,
        iroff3, ksgn, id, jupbnd, k;

    if (limit > QP_LIMIT_MAX) limit = QP_LIMIT_MAX;
    *neval_out = 0; *last_out = 0;
    alist[1] = a; blist[1] = b; rlist[1] = 0.0; elist[1] = 0.0;
    if (epsabs <= 0.0 && epsrel < fmax(50.0 * epmach, 0.5e-28)) {
        *result_out = 0.0; *abserr_out = 0.0;
        return 6;
    }
    /* first approximation to the integral */
    oracle_qk21(a, b, rate, &result, &abserr, &defabs, &resabs);
    dres = fabs(result);
    errbnd = fmax(epsabs, epsrel * dres);
    last = 1;
    rlist[1] = result; elist[1] = abserr; iord[1] = 1;
    if (abserr <= 100.0 * epmach * defabs && abserr > errbnd) ier = 2;
    if (limit == 1) ier = 1;
    if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) goto l140;

    rlist2[1] = result;
    errmax = abserr; maxerr = 1; area = result; errsum = abserr; abserr = oflow;
    nrmax = 1; nres = 0; numrl2 = 2; ktmin = 0; extrap = 0; noext = 0;
    iroff1 = iroff2 = iroff3 = 0;
    ksgn = -1;
    if (dres >= (1.0 - 50.0 * epmach) * defabs) ksgn = 1;

    for (last = 2; last <= limit; ++last) {
        /* bisect the subinterval with the nrmax-th largest error estimate */
        a1 = alist[maxerr];
        b1 = 0.5 * (alist[maxerr] + blist[maxerr]);
        a2 = b1;
        b2 = blist[maxerr];
        erlast = errmax;
        oracle_qk21(a1, b1, rate, &area1, &error1, &resabs, &defab1);
        oracle_qk21(a2, b2, rate, &area2, &error2, &resabs, &defab2);
        area12 = area1 + area2;
        erro12 = error1 + error2;
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
        errbnd = fmax(epsabs, epsrel * fabs(area));
        if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
        if (iroff2 >= 5) ierro = 3;
        if (last == limit) ier = 1;
        if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * epmach) * (fabs(a2) + 1000.0 * uflow)) ier = 4;
        if (error2 > error1) {
            alist[maxerr] = a2; alist[last] = a1; blist[last] = b1;
            rlist[maxerr] = area2; rlist[last] = area1;
            elist[maxerr] = error2; elist[last] = error1;
        } else {
            alist[last] = a2; blist[maxerr] = b1; blist[last] = b2;
            elist[maxerr] = error1; elist[last] = error2;
        }
        qpsrt(limit, last, &maxerr, &errmax, elist, iord, &nrmax);
        if (errsum <= errbnd) goto l115;
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
            /* is the interval to be bisected next the smallest interval? */
            if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
            extrap = 1;
            nrmax = 2;
        }
        if (ierro != 3 && erlarg > ertest) {
            /* smallest interval has the largest error: work down the larger ones first */
            int jumped = 0;
            id = nrmax;
            jupbnd = last;
            if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
            for (k = id; k <= jupbnd; ++k) {
                maxerr = iord[nrmax];
                errmax = elist[maxerr];
                if (fabs(blist[maxerr] - alist[maxerr]) > small) { jumped = 1; break; }
                ++nrmax;
            }
            if (jumped) continue;
        }
        /* perform extrapolation */
        ++numrl2;
        rlist2[numrl2] = area;
        qelg(&numrl2, rlist2, &reseps, &abseps, res3la, &nres);
        ++ktmin;
        if (ktmin > 5 && abserr < 1.0e-3 * errsum) ier = 5;
        if (abseps < abserr) {
            ktmin = 0;
            abserr = abseps;
            result = reseps;
            correc = erlarg;
            ertest = fmax(epsabs, epsrel * fabs(reseps));
            if (abserr <= ertest) break;
        }
        /* prepare bisection of the smallest interval */
        if (numrl2 == 1) noext = 1;
        if (ier == 5) break;
        maxerr = iord[1];
        errmax = elist[maxerr];
        nrmax = 1;
        extrap = 0;
        small *= 0.5;
        erlarg = errsum;
    }
    if (last > limit) last = limit; /* loop ran off the end (only when limit < 2) */

    /* l100: set final result and error estimate */
    if (abserr == oflow) goto l115;
    if (ier + ierro != 0) {
        if (ierro == 3) abserr += correc;
        if (ier == 0) ier = 3;
        if (result != 0.0 && area != 0.0) {
            if (abserr / fabs(result) > errsum / fabs(area)) goto l115;
        } else {
            if (abserr > errsum) goto l115;
            if (area == 0.0) goto l130;
        }
    }
    /* test on divergence */
    if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * 0.01)) {
        if (0.01 > (result / area) || (result / area) > 100.0 || errsum > fabs(area)) ier = 6;
    }
    goto l130;
l115:
    result = 0.0;
    for (k = 1; k <= last; ++k) result += rlist[k];
    abserr = errsum;
l130:
    if (ier > 2) --ier;
l140:
    *neval_out = 42 * last - 21;
    *last_out = last;
    *result_out = result;
    *abserr_out = abserr;
    return ier;
}

/* numpy.isfinite */
static int finite_rate(double r) { return isfinite(r); }

/*
 * PI profile of one locus on the integer grid t = 0..T-1:
 * out[t] = nansum_i townsend_f(t, rates[i])   (compute.py:46-48, bin/tapir_compute.py:119).
 * NaN terms are skipped exactly as numpy.nansum does; summation is sequential in long double
 * (numpy uses pairwise double summation; both differ from any GPU order by O(eps*sqrt(S))).
 */
void oracle_profile(const double *rates, int64_t n_sites, int64_t T, double *out)
{
    for (int64_t t = 0; t < T; ++t) {
        long double acc = 0.0L;
        for (int64_t i = 0; i < n_sites; ++i) {
            double v = townsend_f((double)t, rates[i]);
            if (!isnan(v)) acc += v;
        }
        out[t] = (double)acc;
    }
}

/*
 * Epoch integrals of one locus (compute.py:81-94): for every epoch [a,b] the sequential
 * left-to-right double sum (python builtin sum) over finite-rate sites of quad()'s
 * (integral, abserr).  n_multi counts sites that needed more than one GK21 panel.
 * epoch_ab = {a0,b0,a1,b1,...}.
 */
void oracle_epochs(const double *rates, int64_t n_sites, const double *epoch_ab, int n_epochs,
                   double *sum_integral, double *sum_error, int64_t *n_multi)
{
    for (int e = 0; e < n_epochs; ++e) {
        double si = 0.0, se = 0.0;
        int64_t nm = 0;
        for (int64_t i = 0; i < n_sites; ++i) {
            if (!finite_rate(rates[i])) continue;
            double r, er; int neval, last;
            oracle_qags(epoch_ab[2 * e], epoch_ab[2 * e + 1], rates[i], 1.49e-8, 1.49e-8, 50, &r, &er,
                        &neval, &last);
            si += r; se += er;
            if (last > 1) ++nm;
        }
        sum_integral[e] = si; sum_error[e] = se;
        if (n_multi) n_multi[e] = nm;
    }
}

/* Per-site variant used by the parity tests: result/abserr/last for every site of one epoch. */
void oracle_epoch_sites(const double *rates, int64_t n_sites, double a, double b, double *integral,
                        double *error, int32_t *last_out)
{
    for (int64_t i = 0; i < n_sites; ++i) {
        int neval, last;
        if (!finite_rate(rates[i])) { integral[i] = NAN; error[i] = NAN; last_out[i] = 0; continue; }
        oracle_qags(a, b, rates[i], 1.49e-8, 1.49e-8, 50, &integral[i], &error[i], &neval, &last);
        last_out[i] = last;
    }
}

/*
 * Whole batch in CSR form, the shape the CUDA library consumes: the CPU-port baseline that
 * bench.py times (single thread per call; bench.py fans loci out over processes).
 */
void oracle_batch(int64_t n_loci, const int64_t *offsets, const double *rates, int64_t T,
                  const double *epoch_ab, int n_epochs, double *pi_net /* L*T */,
                  double *sum_integral /* L*E */, double *sum_error /* L*E */)
{
    for (int64_t l = 0; l < n_loci; ++l) {
        const double *r = rates + offsets[l];
        int64_t s = offsets[l + 1] - offsets[l];
        oracle_profile(r, s, T, pi_net + l * T);
        oracle_epochs(r, s, epoch_ab, n_epochs, sum_integral + l * n_epochs, sum_error + l * n_epochs,
                      NULL);
    }
}
