"""GPU tests of the reference-facing Python API (tapir_b200.compute / driver / db): the reference's own
unit tests (tapir/tests/test_compute.py) re-stated against the drop-in module, and the whole
tapir_compute.py --site-rates run of BASELINE config 1 compared with golden rows written by the
reference's db.py."""
import json
import os
import shutil
import sqlite3

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "golden", "ref-data")
RTOL = 1e-9


def rel(a, b, floor=1e-290):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


def test_townsend_vs_phydesign_and_golden(golden):
    """tapir/tests/test_compute.py:41-60 plus the full-precision R fixture the reference never asserts on."""
    from tapir_b200 import compute
    g, meta = golden
    rates = compute.parse_site_rates(os.path.join(DATA, "test-uniform-draw-weights.rates.json"), test=True)
    time = compute.get_time(0, 174)
    townsend = compute.get_townsend_pi(time, np.array(rates))
    assert townsend.shape == (174, 100)
    for k, t in enumerate([10, 20, 50]):
        assert round(abs(np.sum(townsend, axis=1)[t] - meta["phydesign_net_t10_20_50"][k]), 3) == 0
    assert rel(townsend[:101].T, g["fixture/test-R-townsend-output"], floor=1e-300) < 1e-13
    # scalar forms used inside quad(): f(t, rate)
    assert abs(compute.get_townsend_pi(10.0, 0.1) - 16 * 0.01 * 10 * np.exp(-4.0)) < 1e-15
    assert compute.get_townsend_pi(5, np.array([0.0, 0.1])).shape == (2,)
    # NaN propagates exactly as numpy does (culled sites), and nansum skips them
    withnan = np.array(rates)
    withnan[3] = np.nan
    pi = compute.get_townsend_pi(time, withnan)
    assert np.isnan(pi[:, 3]).all() and not np.isnan(np.delete(pi, 3, axis=1)).any()
    picks = compute.get_net_pi_for_periods(pi, [10, 20, -1])
    want = np.nansum(pi, axis=1)
    assert all(abs(picks[t] - want[t]) <= 1e-12 * want[t] for t in (10, 20, -1))
    with pytest.raises(IndexError):
        compute.get_net_pi_for_periods(pi, [174])


def test_integration_vs_phydesign_and_golden(golden):
    """tapir/tests/test_compute.py:62-71 plus test-30-50-integral.npy (per-site, full precision)."""
    from tapir_b200 import compute
    g, meta = golden
    rates = g["bundled_rates"]
    spans = [[0, 10], [10, 15], [15, 20], [20, 30], [20, 70], [20, 100]]
    for k, (a, b) in enumerate(spans):
        out = compute.get_net_integral_for_epochs(rates, [[a, b]])["{0}-{1}".format(a, b)]
        assert round(abs(out["sum(integral)"] - meta["phydesign_integrals"][k]), 4) == 0
        assert rel(out["sum(integral)"], g["unit/sum_integral"][k]) < RTOL
    per_site = np.array([compute.get_integral_over_times(30, 50, r)[0] for r in rates[:10]])
    assert rel(per_site, g["fixture/test-30-50-integral"].reshape(-1)[:10]) < RTOL
    i, e = compute.get_integral_over_times(30, 50, float(rates[0]))
    assert isinstance(i, float) and isinstance(e, float) and e < 1e-8
    with pytest.raises(AssertionError, match="sooner than end time"):
        compute.get_net_integral_for_epochs(rates, [[10, 10]])
    with pytest.raises(ValueError):
        compute.get_net_integral_for_epochs(np.array([]), [[1, 2]])


def test_worker_tuple_matches_reference_worker(golden, tmp_path):
    from tapir_b200 import compute, driver
    g, meta = golden
    work = tmp_path / "locus-one.rates"
    shutil.copyfile(os.path.join(DATA, "test-uniform-draw-weights.rates.json"), work)
    params = [compute.get_time(0, 174), "hyphy2", None, None, str(work), 100, str(work), meta["times"], meta["epochs"], 3, {}]
    name, rates, mean_rate, pi, pi_net, pi_times, pi_epochs = driver.worker(params)
    assert name == str(work) and np.array_equal(rates, g["cfg1/rates"])
    assert rel(mean_rate, g["cfg1/mean_rate"]) < 1e-12
    assert rel(pi, g["cfg1/pi"], floor=1e-300) < 1e-13 and rel(pi_net, g["cfg1/pi_net"]) < RTOL
    assert list(pi_times) == meta["times"] and rel(list(pi_times.values()), g["cfg1/pi_times"]) < RTOL
    labels = ["{0}-{1}".format(a, b) for a, b in meta["epochs"]]
    assert list(pi_epochs) == labels
    assert rel([pi_epochs[k]["sum(integral)"] for k in labels], g["cfg1/integral"]) < RTOL
    assert rel([pi_epochs[k]["sum(error)"] for k in labels], g["cfg1/error"]) < 1e-6
    # all-NaN locus: the reference dies in numpy.vectorize; here a ValueError names the locus
    with pytest.raises(ValueError, match="no finite site rate"):
        compute.compute_loci([np.array([np.nan, np.nan])], compute.get_time(0, 10), [1], [[1, 2]], names=["bad"])


def test_cli_site_rates_end_to_end(golden, tmp_path, capsys):
    """BASELINE config 1: bundled rate JSON (three copies = three loci) + Euteleost.tree through main()."""
    from tapir_b200 import driver
    g, meta = golden
    indir, outdir = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    outdir.mkdir()
    for name in ("uce-1", "uce-2", "uce-3"):
        shutil.copyfile(os.path.join(DATA, "test-uniform-draw-weights.rates.json"), indir / (name + ".rates"))
    rc = driver.main([str(indir), os.path.join(DATA, "Euteleost.tree"), "--output", str(outdir),
                      "--intervals=32-42,88-98,95-105,164-174", "--times=37,93,100,170", "--site-rates", "--multiprocessing"])
    assert rc == 0 and "Storing results" in capsys.readouterr().out
    assert os.path.exists(outdir / "Tree_100_174.0.newick")
    doc = json.load(open(indir / "uce-1.rates"))
    assert len(doc["sites"]["corrected_rates"]) == 100              # side effect of parse_site_rates kept
    conn = sqlite3.connect(str(outdir / "phylogenetic-informativeness.sqlite"))
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_db_golden.json")))
    loci = conn.execute("SELECT id, locus FROM loci ORDER BY id").fetchall()
    assert sorted(n for _, n in loci) == ["uce-1", "uce-2", "uce-3"] and [i for i, _ in loci] == [1, 2, 3]
    assert [r[0] for r in conn.execute("SELECT sql FROM sqlite_master WHERE type='table' ORDER BY name")] == want["ddl"]
    for key, _ in loci:
        net = conn.execute("SELECT time, pi FROM net WHERE id=? ORDER BY time", (key,)).fetchall()
        assert [t for t, _ in net] == list(range(174))
        assert rel([p for _, p in net], [r[2] for r in want["net"]]) < RTOL
        disc = conn.execute("SELECT time, pi FROM discrete WHERE id=? ORDER BY time", (key,)).fetchall()
        assert [t for t, _ in disc] == [r[1] for r in want["discrete"]]
        assert rel([p for _, p in disc], [r[2] for r in want["discrete"]]) < RTOL
        iv = conn.execute("SELECT interval, pi, error FROM interval WHERE id=? ORDER BY interval", (key,)).fetchall()
        assert [r[0] for r in iv] == [r[1] for r in want["interval"]]
        assert rel([r[1] for r in iv], [r[2] for r in want["interval"]]) < RTOL
        assert rel([r[2] for r in iv], [r[3] for r in want["interval"]]) < 1e-6
    # README.rst:178-185: the max-PI epoch of every locus is 164-174, bit-identical to the reference
    from tapir_b200 import db
    assert db.max_pi_epochs(conn) == {"164-174": 3}
    conn.close()


def test_cli_native_ingest_matches_default_path(golden, tmp_path):
    """--no-rate-rewrite (extension): C++ batched ingest, files untouched, same database; subset map honoured."""
    from tapir_b200 import driver
    dbs = []
    for tag, extra in (("a", []), ("b", ["--no-rate-rewrite", "--multiprocessing"])):
        indir, outdir = tmp_path / ("in" + tag), tmp_path / ("out" + tag)
        indir.mkdir()
        outdir.mkdir()
        for name in ("uce-1", "uce-2"):
            shutil.copyfile(os.path.join(DATA, "test-uniform-draw-weights.rates.json"), indir / (name + ".rates"))
        smap = tmp_path / ("map" + tag + ".tsv")
        smap.write_text("uce-2.rates\t10\t60\n")
        before = open(indir / "uce-1.rates").read()
        assert driver.main([str(indir), os.path.join(DATA, "Euteleost.tree"), "--output", str(outdir),
                            "--intervals=32-42,88-98", "--times=37,93", "--site-rates", "--subset-pi-map-file", str(smap)]
                           + extra) == 0
        assert (open(indir / "uce-1.rates").read() == before) == bool(extra)      # rewritten only on the default path
        conn = sqlite3.connect(str(outdir / "phylogenetic-informativeness.sqlite"))
        rows = {}
        for locus, in conn.execute("SELECT locus FROM loci"):
            key = conn.execute("SELECT id FROM loci WHERE locus=?", (locus,)).fetchone()[0]
            rows[locus] = (conn.execute("SELECT time, pi FROM net WHERE id=? ORDER BY time", (key,)).fetchall(),
                           conn.execute("SELECT interval, pi, error FROM interval WHERE id=? ORDER BY interval", (key,)).fetchall())
        conn.close()
        dbs.append(rows)
    assert dbs[0] == dbs[1]                                         # bit-identical results either way
    assert dbs[0]["uce-1"] != dbs[0]["uce-2"]                       # the subset slice changed locus 2


def test_threaded_multi_device_equals_single(golden):
    """--multiprocessing spreads loci over all visible GPUs (a thread and a context per device); with one GPU the
    two 'devices' are the same card, which still exercises partition, per-device contexts and the merge."""
    from tapir_b200 import compute, engine, synthetic
    rng = np.random.default_rng(5)
    loci = [synthetic.uce_rates(rng, int(n), 300, nan_frac=0.03) for n in rng.integers(20, 400, size=57)]
    tv = compute.get_time(0, 300)
    times, epochs = [5, 250], [[32, 42], [88, 98], [164, 174]]
    one = compute.compute_loci(loci, tv, times, epochs, names=list(map(str, range(57))))
    n_dev = engine.device_count()
    many = compute.compute_loci(loci, tv, times, epochs, names=list(map(str, range(57))),
                                devices=[0, 1 % n_dev, 0])
    for a, b in zip(one, many):
        assert a[0] == b[0] and np.array_equal(a[4], b[4]) and a[5] == b[5] and a[6] == b[6]
        np.testing.assert_allclose(a[2], b[2], rtol=0, atol=0, equal_nan=True)
