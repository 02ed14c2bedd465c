"""Pin the CPU oracle against the reference's own golden vectors.

Sources of truth, in order: tests/golden/reference_golden.npz (outputs of the reference's
unmodified tapir/compute.py, see make_golden.py), the reference's unused full-precision
fixtures (test-R-townsend-output.npy, test-30-50-integral.npy) and the PhyDesign 3/4-decimal
pins of tapir/tests/test_compute.py:54-71.
"""
import numpy as np
import pytest

from oracle import capi, townsend_port as tp


def rel(a, b, floor=1e-290):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / (np.abs(b) + floor)) if a.size else 0.0


def test_port_profile_matches_reference_cfg1(golden):
    g, meta = golden
    tcol = tp.time_column(0, 174)
    _, mean_rate, pi, pi_net, pi_times, pi_epochs = tp.locus_worker(tcol, g["cfg1/rates"], meta["times"], meta["epochs"])
    assert np.array_equal(pi, g["cfg1/pi"])
    assert np.array_equal(pi_net, g["cfg1/pi_net"])
    assert float(mean_rate) == float(g["cfg1/mean_rate"])
    assert [float(pi_times[t]) for t in meta["times"]] == list(g["cfg1/pi_times"])
    labels = ["{0}-{1}".format(a, b) for a, b in meta["epochs"]]
    assert [float(pi_epochs[k]["sum(integral)"]) for k in labels] == list(g["cfg1/integral"])
    assert [float(pi_epochs[k]["sum(error)"]) for k in labels] == list(g["cfg1/error"])
    assert tp.argmax_epoch(pi_epochs, meta["epochs"]) == 3  # '164-174', BASELINE.md section 3


def test_survey_a5_table(golden):
    """SURVEY.md A.5 / BASELINE.md section 3 literal values."""
    g, _ = golden
    assert g["cfg1/pi_net"][173] == 0.0013947599572823488 and int(np.argmax(g["cfg1/pi_net"])) == 173
    np.testing.assert_array_equal(g["cfg1/pi_times"], [0.0005125287881276899, 0.0010307012901663213,
                                                        0.0010778251204444581, 0.0013870113425963346])
    np.testing.assert_array_equal(g["cfg1/integral"], [0.005121028310508198, 0.010304019918297151,
                                                       0.010775390910162838, 0.013841636570348266])
    np.testing.assert_array_equal(g["cfg1/error"], [5.685483540083456e-17, 1.1439760159489734e-16,
                                                    1.1963087087799403e-16, 1.5367303618894218e-16])


def test_reference_fixture_R_townsend(golden):
    """tests/test-data/test-R-townsend-output.npy: (100 sites, t=0..100), unused by the reference's tests."""
    g, _ = golden
    fx = g["fixture/test-R-townsend-output"]
    pi = tp.pi_matrix(tp.time_column(0, 101), g["bundled_rates"])
    assert fx.shape == (100, 101)
    assert rel(pi.T, fx, floor=1e-30) < 1e-14
    c = np.array([[capi.lib().oracle_townsend_pi(float(t), float(r)) for t in range(101)] for r in g["bundled_rates"]])
    assert rel(c, fx, floor=1e-30) < 1e-14


def test_reference_fixture_30_50_integral(golden):
    g, _ = golden
    fx = g["fixture/test-30-50-integral"].reshape(-1)
    integ, err, last = capi.epoch_sites(g["bundled_rates"], 30, 50)
    assert rel(integ, fx) < 1e-14
    assert abs(integ.sum() - 0.37723037831767525) < 1e-15
    vec = np.vectorize(tp.site_integral)
    i2, _ = vec(30, 50, g["bundled_rates"])
    assert rel(i2, fx) < 1e-14


def test_phydesign_pins(golden):
    """tapir/tests/test_compute.py:54-71 (3 and 4 decimal places)."""
    g, meta = golden
    rates = g["bundled_rates"]
    net = capi.profile(rates, 174)
    for t, want in zip([10, 20, 50], meta["phydesign_net_t10_20_50"]):
        assert round(abs(net[t] - want), 3) == 0
    for (a, b), want in zip([[0, 10], [10, 15], [15, 20], [20, 30], [20, 70], [20, 100]], meta["phydesign_integrals"]):
        si, _, _ = capi.epochs(rates, [[a, b]])
        assert round(abs(si[0] - want), 4) == 0


def test_c_port_vs_reference_per_site_integrals(golden):
    """C QAGS vs the reference's per-site quad() results, incl. [20,100] where one site is multi-panel."""
    g, _ = golden
    for k, (a, b) in enumerate(g["unit/spans"]):
        integ, err, last = capi.epoch_sites(g["bundled_rates"], a, b)
        # numpy's exp / pow differ from libm's by <= 1 ulp per integrand call: not bit-exact, but 1e-13
        assert rel(integ, g["unit/site_integral"][k]) < 1e-13
        np.testing.assert_allclose(err, g["unit/site_error"][k], rtol=1e-6, atol=1e-300)


@pytest.mark.parametrize("tag", ["synA", "synB"])
def test_c_port_vs_reference_synthetic(golden, tag):
    g, _ = golden
    off, rates, T = g[tag + "/offsets"], g[tag + "/rates"], int(g[tag + "/T"])
    pi_net, si, se = capi.batch(off, rates, T, g[tag + "/epochs"])
    assert rel(pi_net, g[tag + "/pi_net"]) < 1e-12
    assert rel(si, g[tag + "/integral"]) < 1e-12
    np.testing.assert_allclose(se, g[tag + "/error"], rtol=1e-6, atol=1e-300)
    assert np.array_equal(np.argmax(si, axis=1), np.argmax(g[tag + "/integral"], axis=1))


def test_port_vs_reference_synthetic_bitexact(golden):
    """The numpy/scipy port reproduces the reference's outputs exactly (same third-party arithmetic)."""
    g, _ = golden
    tag = "synB"
    off, rates, T = g[tag + "/offsets"], g[tag + "/rates"], int(g[tag + "/T"])
    tcol = tp.time_column(0, T)
    epochs = [list(map(int, e)) for e in g[tag + "/epochs"]]
    for l in range(2):
        lam = rates[off[l]:off[l + 1]]
        _, mean_rate, _, pi_net, _, pi_epochs = tp.locus_worker(tcol, lam, [0], epochs, keep_matrix=False)
        assert np.array_equal(pi_net, g[tag + "/pi_net"][l])
        labels = ["{0}-{1}".format(a, b) for a, b in epochs]
        assert [float(pi_epochs[k]["sum(integral)"]) for k in labels] == list(g[tag + "/integral"][l])


def test_tree_rule(golden):
    _, meta = golden
    assert tp.tree_depth_and_correction(meta["tree"]) == (174.0, 100)
