"""Pin oracle/quadpack_port.c bit-for-bit against scipy.integrate.quad (the reference's
third-party integrator, compute.py:52) by driving scipy with the oracle's own integrand."""
import warnings

import numpy as np
import pytest
from scipy import integrate

from oracle import capi

SPANS = [(32, 42), (88, 98), (95, 105), (164, 174), (0, 10), (20, 70), (20, 100), (0, 174), (0, 1000), (3, 4000)]


@pytest.mark.parametrize("span", SPANS)
def test_qags_bit_exact(span):
    f = capi.lib().oracle_townsend_pi
    a, b = span
    n_multi = 0
    rates = np.concatenate([np.logspace(-7, 2.5, 160), [0.0, 0.1, 0.5, 1.0, 2.0, 3.3]])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in rates:
            r = float(r)
            out = integrate.quad(lambda t, rr: f(t, rr), a, b, args=(r,), full_output=1)
            res, err, neval, last, ier = capi.qags(float(a), float(b), r)
            assert (res, err, neval, last) == (out[0], out[1], out[2]["neval"], out[2]["last"]), (span, r)
            assert (ier != 0) == (len(out) > 3)
            n_multi += last > 1
    if span in [(0, 10), (0, 174), (0, 1000), (3, 4000)]:
        assert n_multi > 5  # the bisection / epsilon-extrapolation path really ran


def test_qk21_single_panel_formula():
    res, err, resabs, resasc = capi.qk21(32.0, 42.0, 1e-3)
    # f >= 0 on t >= 0: resabs accumulates the same terms in the same order as resk
    assert resabs == abs(res)
    assert err == 50 * np.finfo(float).eps * resabs
