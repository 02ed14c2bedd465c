"""CPU checks of the algebra the CUDA kernels rely on (no GPU): the constant tables of gk21.cuh are re-derived in
high-precision decimal arithmetic, and the moment form of profile_mom_kernel is restated in numpy and compared with
the direct evaluation at the slow/fast threshold, where its series remainder is largest."""
import os
import re
from decimal import Decimal as D, getcontext
from math import factorial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
getcontext().prec = 80


def _array(name, text):
    m = re.search(r"c_%s\[\d+\]\s*=\s*\{(.*?)\};" % name, text, re.S)
    assert m, name
    return [float(v) for v in re.findall(r"[-+]?\d\.?\d*(?:[eE][-+]?\d+)?", m.group(1))]


def _rule():
    text = open(os.path.join(ROOT, "tapir_b200", "csrc", "gk21.cuh")).read()
    c = {n: _array(n, text) for n in ("xgk", "wgk", "wg", "wx", "dl", "dlx", "ak", "bk")}
    assert [len(c[n]) for n in ("xgk", "wgk", "wg", "wx", "dl", "dlx", "ak", "bk")] == [10, 11, 5, 10, 10, 10, 7, 7]
    return c


def test_gk21_product_constants():
    c = _rule()
    x, w, g = [D(v) for v in c["xgk"]], [D(v) for v in c["wgk"]], [D(v) for v in c["wg"]]
    for j in range(10):
        dl = w[j] - (g[(j - 1) // 2] if j % 2 == 1 else 0)
        assert c["wx"][j] == float(w[j] * x[j])
        assert c["dl"][j] == float(dl)
        assert c["dlx"][j] == float(dl * x[j])
    # Kronrod and Gauss both integrate constants exactly: the difference weights sum to zero
    assert abs(float(2 * sum(D(v) for v in c["dl"]) + w[10])) < 1e-16


def test_gk21_series_coefficients():
    """c_ak / c_bk: moments of the Kronrod rule; seven terms reproduce the pair sums below y x_0 log2(e) = 1/2."""
    c = _rule()
    x, w = [D(v) for v in c["xgk"]], [D(v) for v in c["wgk"]]
    for n in range(7):
        ak = (2 * sum(w[j] * x[j] ** (2 * n) for j in range(10)) + (w[10] if n == 0 else 0)) / factorial(2 * n)
        bk = 2 * sum(w[j] * x[j] ** (2 * n + 2) for j in range(10)) / factorial(2 * n + 1)
        assert c["ak"][n] == float(ak) and c["bk"][n] == float(bk)
    for y in (0.348, 0.1, 1e-3):
        yd = D(y)
        direct_a = w[10] + sum(w[j] * ((yd * x[j]).exp() + (-yd * x[j]).exp()) for j in range(10))
        direct_b = sum(w[j] * x[j] * ((yd * x[j]).exp() - (-yd * x[j]).exp()) for j in range(10))
        sa = sum(D(c["ak"][n]) * yd ** (2 * n) for n in range(7))
        sb = sum(D(c["bk"][n]) * yd ** (2 * n + 1) for n in range(7))
        assert abs(sa - direct_a) / direct_a < D("3e-16") and abs(sb - direct_b) / direct_b < D("3e-16")


def _moment_form(rates, T0, TT=1024):
    """numpy restatement of profile_mom_kernel's arithmetic for one tile: sum_s a_s q_s^t for t = T0 .. T0 + TT - 1."""
    lam = np.asarray(rates, dtype=np.float64)
    a, c = 16.0 * lam * lam, 4.0 * lam
    u0 = a * np.exp(-(c * float(T0)))
    z = -64.0 * c
    q64 = np.exp(z)
    nb = TT // 128
    m = np.empty((32, lam.size))
    z2 = z * z
    z4 = z2 * z2
    m[0], m[1], m[2], m[3] = u0, u0 * z, u0 * z2, (u0 * z) * z2
    for k in range(4, 32):
        m[k] = m[k - 4] * z4
    w = np.empty((nb, lam.size))
    w[0] = q64
    q128 = q64 * q64
    for hb in range(1, nb):
        w[hb] = w[hb - 1] * q128
    M = m @ w.T                                   # [32][nb], the contraction over sites
    t = np.arange(128, dtype=np.float64)
    x = (t - 64.0) / 64.0
    P = np.empty((128, 32))
    for k in range(32):
        P[:, k] = x ** k / float(factorial(k))
    return (P @ M).T.reshape(-1)                  # [128 hb + t]


def test_moment_form_remainder_at_the_threshold():
    rng = np.random.default_rng(3)
    thr = float(np.uint64(0x3F94000000000000).view(np.float64))
    assert thr == 0.01953125                      # rate_is_slow: 64 c = 256 r <= 5
    for rates in (rng.uniform(0.0190, thr, 64), rng.gamma(0.5, 0.004, 400).clip(1e-9, thr * 0.999),
                  np.array([1e-7, 1e-5, thr * 0.9999])):
        for T0 in (0, 1024, 4096):
            got = _moment_form(rates, T0)
            t = T0 + np.arange(1024, dtype=np.float64)
            want = (16.0 * rates[:, None] ** 2 * np.exp(-4.0 * rates[:, None] * t[None, :])).sum(axis=0)
            err = np.max(np.abs(got - want) / want)
            assert err < 5e-11, (T0, err)         # remainder 5^32/32! against e^-5 = 1.3e-11


def _collapsed_panel(c, lam, a_lim, b_lim, series):
    """numpy restatement of quad_kernel's collapsed GK21 panel (gk21_collapsed + panel_sums[_series]) for one site and
    one epoch -> (result, abserr, proved)."""
    eps, uflow = np.finfo(float).eps, np.finfo(float).tiny
    x, w = np.array(c["xgk"]), np.array(c["wgk"])
    fc, h = 0.5 * (a_lim + b_lim), 0.5 * (b_lim - a_lim)
    a, cc = 16.0 * lam * lam, 4.0 * lam
    y = cc * h
    R, Dn = np.exp(y * x), np.exp(-y * x)
    if series:
        s = y * y
        pa, pb = c["ak"][6], c["bk"][6]
        for n in range(5, -1, -1):
            pa, pb = pa * s + c["ak"][n], pb * s + c["bk"][n]
        Ak, Bk, dA, dB, e0rel = pa, (pb * y) * h, 0.0, 0.0, 32.0 * eps
    else:
        S, Dl = R + Dn, R - Dn
        Ak = w[10] + np.dot(w[:10], S)
        Bk = np.dot(np.array(c["wx"]), Dl) * h
        dA = w[10] + np.dot(np.array(c["dl"]), S)
        dB = np.dot(np.array(c["dlx"]), Dl) * h
        e0rel = 0.0
    scale = (a * np.exp(-(cc * fc))) * h
    resk = fc * Ak - Bk
    e0 = e0rel * resk + abs(fc * dA - dB)
    reskh = 0.5 * resk
    n, fl = 200.0 * e0 / resk, 50.0 * eps
    f14, f24 = (fc - h * x[4]) * R[4], (fc + h * x[4]) * Dn[4]
    f15, f25 = (fc - h * x[5]) * R[5], (fc + h * x[5]) * Dn[5]
    lb = (w[4] * (abs(f14 - reskh) + abs(f24 - reskh)) + w[5] * (abs(f15 - reskh) + abs(f25 - reskh))) / resk
    res = resk * scale
    proved = bool(res > uflow / (50.0 * eps) and n < lb and n ** 3 <= fl * fl * lb and lb > fl)
    return res, 50.0 * eps * res, proved


def test_collapsed_panel_against_quadpack():
    """Whenever the floor proof of the collapsed panel holds, QUADPACK's dqk21 (oracle C port) returns the same
    integral to rounding and exactly the rounding floor as its error estimate -- on both table branches."""
    from oracle import capi
    c = _rule()
    eps = np.finfo(float).eps
    rng = np.random.default_rng(17)
    n_proved = n_unproved = 0
    epochs = [(32, 42), (88, 98), (164, 174), (0, 10), (5, 45), (10, 110), (1, 2)]
    for lam in np.concatenate([rng.gamma(0.5, 0.004, 300), rng.uniform(0.01, 0.5, 150), [1e-6, 1e-4, 2.5e-3]]):
        if lam <= 0.0:
            continue
        for (a_lim, b_lim) in epochs:
            y = 4.0 * lam * 0.5 * (b_lim - a_lim)
            for series in ((True, False) if y * 1.4365 < 0.4999 else (False,)):
                res, err, proved = _collapsed_panel(c, lam, a_lim, b_lim, series)
                q_res, q_err, q_resabs, q_resasc = capi.qk21(float(a_lim), float(b_lim), float(lam))
                if proved:
                    n_proved += 1
                    assert abs(res - q_res) <= 1e-13 * abs(q_res), (lam, a_lim, b_lim, series)
                    assert abs(q_err - 50.0 * eps * q_resabs) <= 1e-12 * q_err, (lam, a_lim, b_lim, series, q_err)
                    assert q_err != q_resasc
                else:
                    n_unproved += 1
    assert n_proved > 2000 and n_unproved > 100, (n_proved, n_unproved)


def test_inverse_factorial_table():
    """c_inv_fact of kernels.cuh (the 1/k! that profile_mom_kernel puts onto the moments) is 1/k! correctly rounded."""
    text = open(os.path.join(ROOT, "tapir_b200", "csrc", "kernels.cuh")).read()
    vals = _array("inv_fact", text)
    assert len(vals) == 32
    for k, v in enumerate(vals):
        assert v == float(D(1) / D(factorial(k))), k


def _table(name, text):
    m = re.search(r"%s\[[^\]]*\]\s*=\s*\{(.*?)\};" % name, text, re.S)
    assert m, name
    return [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))]


def test_shift_tables_of_the_moment_classes():
    """g_tile_shift[n][j] = (2 j - 7)^n / n! (whole tile -> its eight parts), correctly rounded; g_inv_fact is the
    constant-bank table again; both as written in kernels.cuh."""
    from fractions import Fraction
    text = open(os.path.join(ROOT, "tapir_b200", "csrc", "kernels.cuh")).read()
    shift = _table("g_tile_shift", text)
    assert len(shift) == 128
    for n in range(16):
        for j in range(8):
            assert shift[8 * n + j] == float(Fraction((2 * j - 7) ** n, factorial(n))), (n, j)
    inv_c, inv_g = _table("c_inv_fact", text), _table("g_inv_fact", text)
    assert len(inv_c) == 32 and inv_c == inv_g
    for k in range(32):
        assert inv_c[k] == float(Fraction(1, factorial(k)))


def test_block_shifts_are_changes_of_variable():
    """The shifts that move the coarser classes' sums onto the 128-point blocks (profile_mom_kernel):
    M[m] = sum_{k >= m} s^-k o^(k-m)/(k-m)! N[k] re-expands sum_k N[k]/k! y^k in x, y = (x + o)/s -- exactly, for any
    polynomial; with the moments of a class at its bound the re-expanded profile is the class's own truncation."""
    rng = np.random.default_rng(5)
    for s, offsets, terms in ((2, (-1.0, 1.0), 16), (2, (-1.0, 1.0), 32), (8, tuple(2.0 * j - 7.0 for j in range(8)), 16)):
        N = rng.standard_normal(terms) * np.array([float(factorial(k)) ** 0.5 for k in range(terms)])
        for o in offsets:
            M = [sum(float(s) ** -k * o ** (k - m) / factorial(k - m) * N[k] for k in range(m, terms)) for m in range(terms)]
            x = np.linspace(-1.0, 1.0, 41)
            y = (x + o) / s
            want = sum(N[k] / factorial(k) * y ** k for k in range(terms))
            got = sum(M[m] / factorial(m) * x ** m for m in range(terms))
            assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)) + 1e-13
    # a site at the whole-tile bound of a 2,048-point tile ((TT/2) c = 1.25): tile -> 256-point parts -> 128-point blocks
    lam = 0.0048828125 * 128.0 / 2048.0 * (1 - 1e-12)
    a, c = 16.0 * lam * lam, 4.0 * lam
    S = np.array([a * np.exp(-c * 1024.0) * (-1024.0 * c) ** k for k in range(16)])
    worst = 0.0
    for hb in range(16):
        j, half = hb // 2, (1.0 if hb % 2 else -1.0)
        Nc = [sum(8.0 ** -k * (2.0 * j - 7.0) ** (k - m) / factorial(k - m) * S[k] for k in range(m, 16)) for m in range(16)]
        Mf = [sum(2.0 ** -k * half ** (k - m) / factorial(k - m) * Nc[k] for k in range(m, 16)) for m in range(16)]
        t = 128.0 * hb + np.arange(128.0)
        x = (t - (128.0 * hb + 64.0)) / 64.0
        got = sum(Mf[m] / factorial(m) * x ** m for m in range(16))
        want = a * np.exp(-c * t)
        worst = max(worst, float(np.max(np.abs(got - want) / want)))
    assert worst < 6e-12, worst                     # 1.25^16 / 16! = 1.7e-12 against e^-1.25
