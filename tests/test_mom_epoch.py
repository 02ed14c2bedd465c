"""CPU checks of the moment-functional epoch integrals (tapir_b200/csrc/mom_epoch.h), through the library's own
host-side export tapir_mom_epoch_plan -- the coefficients and rate windows tested here are the shipped ones.

For the sites the plan declares served by the moments, two things must hold against the QUADPACK port
(oracle/quadpack_port.c, pinned bit-for-bit to scipy.integrate.quad):
  (1) soundness: QAGS returns the first GK21 panel with the rounding floor 50 eps resabs as its error estimate
      (so sum(error) = 50 eps sum(integral), and no site is bisected);
  (2) value: sum_j p[j] M[j][hb], with the moments formed as profile_mom_kernel forms them, equals the sum of the
      per-site integrals far inside the 1e-9 gate.
No GPU is used; the kernels' use of the plan is covered by tests/test_gpu_parity.py.
"""
import ctypes
import math

import numpy as np
import pytest

from oracle import capi
from tapir_b200 import _lib

EPS = np.finfo(float).eps
CLASS_BOUND = 0.01953125


BOUND = {32: 0.01953125, 16: 0.0048828125}       # the moment classes of profile_mom_kernel<RH, KM>


def plan(a, b, n_tiles=8, tile_points=1024, km=32, class_x=0.0):
    out = _lib.MomEpoch()
    _lib.check(_lib.load().tapir_mom_epoch_plan(float(a), float(b), n_tiles, tile_points, km, class_x, ctypes.byref(out)))
    out.km = km
    out.bound = class_x / 256.0 if class_x else BOUND[km]
    return out


def moments(rates, tile, hb, tile_points, km=32, bound=None, one_level=False):
    """M[k] = sum_s a e^{-c T0} (-64 c)^k e^{-c (128 hb + 64)} of block hb of the tile -- formed as profile_mom_kernel
    forms it (kernels.cuh, moment pass): sites below 0.0048828125 x 128 / tile_points are one 16-term polynomial on the
    whole tile, shifted onto its eight parts; on 2,048-point tiles the sites below half the class bound are expanded on
    256-point blocks (z = -128 c) and those moments, with the whole-tile part, shifted onto the 128-point blocks:
    M[m] += sum_{k >= m} s^-k o^(k-m) / (k-m)! N[k]  for a part at offset o of a block s times as long."""
    lam = np.asarray(rates, dtype=np.float64)
    bound = BOUND[km] if bound is None else bound
    whole = lam < (0.0 if one_level else 0.0048828125 * 128.0 / tile_points)      # one_level: profile_f32_kernel
    coarse = (lam < 0.5 * bound) & ~whole if tile_points == 2048 and not one_level else np.zeros(lam.shape, dtype=bool)

    def block_moments(lam, half, blk, terms):
        a, c = 16.0 * lam * lam, 4.0 * lam
        u0 = a * np.exp(-(c * float(tile * tile_points)))
        z = -half * c
        w = np.exp(z) * np.exp(2.0 * z) ** blk       # q^half (q^(2 half))^blk
        return np.array([np.sum((u0 * z ** k) * w) if k < terms else 0.0 for k in range(32)])

    def shift(N, scale, o, terms):
        return np.array([sum(float(scale) ** -k * o ** (k - m) / math.factorial(k - m) * N[k] for k in range(m, terms))
                         if m < terms else 0.0 for m in range(32)])

    M = block_moments(lam[~coarse & ~whole], 64.0, hb, km)
    if tile_points == 2048:
        N = block_moments(lam[coarse], 128.0, hb // 2, km)
        if whole.any():
            N = N + shift(block_moments(lam[whole], 1024.0, 0, 16), 8, 2.0 * (hb // 2) - 7.0, 16)
        M = M + shift(N, 2, 1.0 if hb % 2 else -1.0, km)
    elif whole.any():
        M = M + shift(block_moments(lam[whole], tile_points / 2.0, 0, 16), 8, 2.0 * hb - 7.0, 16)
    return M


def served(pl, rates):
    r = np.asarray(rates)
    ok = (r > 1.57e-162) & (r < getattr(pl, "bound", BOUND[32])) & (r >= pl.r_min) & (r <= pl.r_max)
    return ok & ~((r >= pl.w_lo) & (r <= pl.w_hi))


EPOCHS = [(32, 42), (88, 98), (95, 105), (164, 174), (0, 10), (0, 1), (3, 40), (5, 60), (0, 64), (60, 68), (120, 136),
          (127, 129), (500, 564), (1000, 1002), (1023, 1025), (2100, 2110), (4000, 4050), (7000, 7010)]


@pytest.mark.parametrize("km", [32, 16])
@pytest.mark.parametrize("span", EPOCHS)
def test_functional_matches_quadpack_sums(span, km):
    a, b = span
    for tile_points in (1024, 2048):
        pl = plan(a, b, n_tiles=8, tile_points=tile_points, km=km)
        assert pl.eligible, span
        C = 0.5 * (a + b)
        gb = int(C // 128)
        assert (pl.tile, pl.hb, pl.tb) == (gb // (tile_points // 128), gb % (tile_points // 128), 128.0 * gb + 64.0)
        rng = np.random.default_rng(a * 1000 + b)
        lam = np.concatenate([rng.gamma(0.5, 0.004, 3000), rng.uniform(0, BOUND[km], 1000),
                              10.0 ** rng.uniform(-15, -3, 500)])
        lam = lam[served(pl, lam)]
        assert lam.size > 300 and np.all(np.array(pl.p[km:]) == 0.0)
        integ, err, last = capi.epoch_sites(lam, a, b)
        assert np.all(last == 1)
        want = float(np.sum(integ))
        got = float(np.dot(np.array(pl.p[:]), moments(lam, pl.tile, pl.hb, tile_points, km)))
        assert abs(got - want) <= 2e-11 * want, (span, got, want)
        # the error column: every served site carries the rounding floor
        np.testing.assert_allclose(np.sum(err), 50 * EPS * want, rtol=1e-12)


@pytest.mark.parametrize("km", [32, 16])
@pytest.mark.parametrize("span", EPOCHS)
def test_served_sites_get_the_rounding_floor(span, km):
    """Dense sweep of the served rate range, with extra points at the edges of the excluded window, at the class
    bound and at r_max: dqk21's estimate is exactly 50 eps resabs, dqagse accepts the first panel."""
    a, b = span
    pl = plan(a, b, km=km)
    edges = [pl.r_min, pl.r_max, pl.w_lo, pl.w_hi, BOUND[km]]
    near = np.concatenate([e * (1.0 + np.linspace(-3e-3, 3e-3, 601)) for e in edges if e > 0])
    C = 0.5 * (a + b)
    peak = np.linspace(0.9, 1.1, 4001) / (4.0 * C)             # sites whose PI curve peaks inside the epoch
    lam = np.concatenate([np.geomspace(2.5e-16, BOUND[km], 20000), near, peak,
                          np.nextafter(pl.w_lo, 0.0, dtype=np.float64).reshape(1),
                          np.nextafter(pl.w_hi, 1.0, dtype=np.float64).reshape(1)])
    lam = np.unique(lam[served(pl, lam)])
    assert lam.size > 10000
    integ, err, last = capi.epoch_sites(lam, a, b)
    assert np.all(last == 1)
    floor = np.array([50 * EPS * capi.qk21(a, b, r)[2] for r in lam[::97]])
    assert np.array_equal(err[::97], floor)
    # resabs = result for a non-negative integrand, up to the rounding of two different sums
    np.testing.assert_allclose(err, 50 * EPS * integ, rtol=1e-13)


def test_window_sits_on_the_peak_and_is_narrow():
    for (a, b), km in [(e, k) for e in EPOCHS for k in (32, 16)]:
        pl = plan(a, b, km=km)
        C, h = 0.5 * (a + b), 0.5 * (b - a)
        if pl.w_lo <= pl.w_hi:
            # the window brackets the rate whose curve t e^{-4 r t} peaks at the epoch centre, if that rate is in range
            if 1.0 / (4 * C) < pl.r_max:
                assert pl.w_lo < 1.0 / (4 * C) * (1 + (h / C) ** 2) and pl.w_hi > 1.0 / (4 * C)
            assert (pl.w_hi - pl.w_lo) <= 0.05 * pl.w_hi + 1e-300, (a, b, pl.w_lo, pl.w_hi)
        assert pl.r_max <= BOUND[km] and pl.r_max * 4 * h <= 1.0 + 1e-12


@pytest.mark.parametrize("span,n_tiles,why", [
    ((-1, 5), 8, "negative limit"), ((0, 100), 8, "half-length above 32"), ((0, 0.5), 8, "half-length below 1/2"),
    ((9000, 9010), 8, "centre beyond the grid's tiles"), ((1020, 1030), 1, "centre beyond a one-tile grid"),
    ((0.1, 10.3), 8, "centre not representable exactly"),
])
def test_ineligible_epochs(span, n_tiles, why):
    pl = plan(span[0], span[1], n_tiles=n_tiles)
    assert not pl.eligible, why


def test_truncation_at_the_corner_of_the_range():
    """Largest remainder: a site at r_max of an epoch whose centre sits at the edge of its block."""
    for (a, b), km in [(e, k) for e in [(120, 136), (0, 64), (250, 262), (1, 127)] for k in (32, 16)]:
        pl = plan(a, b, km=km)
        if not pl.eligible:
            continue
        r = np.nextafter(pl.r_max, 0.0)
        if not served(pl, np.array([r]))[0]:
            continue
        integ, _, last = capi.epoch_sites(np.array([r]), a, b)
        got = float(np.dot(np.array(pl.p[:]), moments(np.array([r]), pl.tile, pl.hb, 1024, km)))
        assert last[0] == 1 and abs(got - integ[0]) <= 5e-11 * integ[0], (a, b, r, got, integ[0])


@pytest.mark.parametrize("span", EPOCHS)
def test_fp32_mode_class(span):
    """TAPIR_FP32_MODE: 16 moments with the class 64 c <= 2.5.  Soundness is unchanged (QUADPACK returns the first panel
    with the rounding floor for every served site); the value is inside that mode's 1e-5 gate by an order of magnitude."""
    a, b = span
    pl = plan(a, b, km=16, class_x=2.5)
    assert pl.eligible and pl.r_max <= 2.5 / 256
    rng = np.random.default_rng(a + 7 * b)
    lam = np.concatenate([rng.gamma(0.5, 0.004, 3000), rng.uniform(0, 2.5 / 256, 2000), [np.nextafter(pl.r_max, 0.0)]])
    lam = lam[served(pl, lam)]
    assert lam.size > 300
    integ, err, last = capi.epoch_sites(lam, a, b)
    assert np.all(last == 1)
    np.testing.assert_allclose(err, 50 * EPS * integ, rtol=1e-13)
    got = float(np.dot(np.array(pl.p[:]), moments(lam, pl.tile, pl.hb, 1024, 16, bound=2.5 / 256, one_level=True)))
    assert abs(got - np.sum(integ)) <= 2e-6 * np.sum(integ)
    worst = np.array([np.nextafter(pl.r_max, 0.0)])
    if served(pl, worst)[0]:
        one = float(np.dot(np.array(pl.p[:]), moments(worst, pl.tile, pl.hb, 1024, 16, bound=2.5 / 256, one_level=True)))
        assert abs(one - capi.epoch_sites(worst, a, b)[0][0]) <= 5e-6 * one
