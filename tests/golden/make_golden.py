#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF.  Run in the authoring container only
(/root/reference does not exist on the GPU box); the produced fixtures are committed.

    python tests/golden/make_golden.py

How the reference is executed: /root/reference/tapir/compute.py is imported UNMODIFIED under
Python 3 with two shims: a stub `dendropy` module (absent here, only used by
the tree / alignment functions, which are not called) and `scipy.vectorize = numpy.vectorize`
(removed from modern scipy).  The per-locus driver below follows bin/tapir_compute.py:104-123
line by line (that script is Python-2 only and cannot be imported).

Also bundles the reference's own fixtures that pin this path
(tapir/tests/test-data/*.npy, the rate JSON's 100 rates, Euteleost.tree) so the tests on the
GPU box need nothing outside the repo.
"""
import importlib.util
import json
import os
import sys
import types

import numpy
import scipy

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
TD = os.path.join(REF, "tapir", "tests", "test-data")

EPOCHS = [[32, 42], [88, 98], [95, 105], [164, 174]]
TIMES = [37, 93, 100, 170]


def load_reference_compute():
    sys.modules.setdefault("dendropy", types.ModuleType("dendropy"))
    scipy.vectorize = numpy.vectorize
    spec = importlib.util.spec_from_file_location("ref_compute", os.path.join(REF, "tapir", "compute.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_worker(ref, time_vector, rates, times, epochs):
    """bin/tapir_compute.py:110-122 with the reference's own functions."""
    mean_rate = numpy.mean(numpy.ma.masked_array(rates, numpy.isnan(rates)))
    pi = ref.get_townsend_pi(time_vector, rates)
    pi_net = numpy.nansum(pi, axis=1)
    pi_times = ref.get_net_pi_for_periods(pi, times)
    pi_epochs = ref.get_net_integral_for_epochs(rates[numpy.isfinite(rates)], epochs)
    return mean_rate, pi, pi_net, pi_times, pi_epochs


def epochs_to_arrays(pi_epochs, epochs):
    keys = ["{0}-{1}".format(a, b) for a, b in epochs]
    return (numpy.array([float(pi_epochs[k]["sum(integral)"]) for k in keys]),
            numpy.array([float(pi_epochs[k]["sum(error)"]) for k in keys]))


def synthetic_loci(seed, n_loci, smin, smax, T, nan_frac=0.05, zero_frac=0.3, fast_frac=0.0):
    """UCE-like rates: gamma(0.5) with mean ~2/T, exact zeros, NaN (culled) sites."""
    rng = numpy.random.default_rng(seed)
    loci = []
    for _ in range(n_loci):
        s = int(rng.integers(smin, smax + 1))
        lam = rng.gamma(0.5, 4.0 / T, size=s)
        lam[rng.random(s) < zero_frac] = 0.0
        if fast_frac:
            fast = rng.random(s) < fast_frac
            lam[fast] = rng.uniform(0.05, 3.0, size=int(fast.sum()))
        lam[rng.random(s) < nan_frac] = numpy.nan
        loci.append(lam)
    return loci


class Py2Dict(dict):
    iteritems = dict.items


def main():
    ref = load_reference_compute()
    out = {}
    rate_json = os.path.join(TD, "test-uniform-draw-weights.rates.json")
    raw = ref.parse_site_rates(rate_json, test=True)
    out["bundled_rates"] = raw
    out["bundled_rates_div10"] = ref.parse_site_rates(rate_json, 10., test=True)
    for name in ["test-parsed-rates", "test-R-townsend-output", "test-30-50-integral",
                 "test-culled-rates", "chr1_918-test-cutoff-values"]:
        out["fixture/" + name] = numpy.load(os.path.join(TD, name + ".npy"))
    # --- config 1: bundled data, cf=100, T=174 (bin/tapir_compute.py:133-135)
    cf, depth = 100, 174.0
    tv = ref.get_time(0, int(depth))
    rates = ref.parse_site_rates(rate_json, cf, test=True)
    mean_rate, pi, pi_net, pi_times, pi_epochs = ref_worker(ref, tv, rates, TIMES, EPOCHS)
    cfg1 = (rates, mean_rate, pi, pi_net, Py2Dict(pi_times), Py2Dict(pi_epochs))
    out["cfg1/rates"] = rates
    out["cfg1/mean_rate"] = numpy.float64(mean_rate)
    out["cfg1/pi"] = pi
    out["cfg1/pi_net"] = pi_net
    out["cfg1/pi_times"] = numpy.array([float(pi_times[t]) for t in TIMES])
    out["cfg1/integral"], out["cfg1/error"] = epochs_to_arrays(pi_epochs, EPOCHS)
    # --- the unit-test shape: cf=1, T=174 (tapir/tests/test_compute.py:41-71)
    tv = ref.get_time(0, 174)
    pi1 = ref.get_townsend_pi(tv, numpy.array(raw))
    out["unit/pi_net_sum"] = numpy.sum(pi1, axis=1)
    spans = [[0, 10], [10, 15], [15, 20], [20, 30], [20, 70], [20, 100], [30, 50]]
    vec = numpy.vectorize(ref.get_integral_over_times)
    out["unit/spans"] = numpy.array(spans, dtype=numpy.int64)
    ints, errs = [], []
    for a, b in spans:
        i, e = vec(a, b, raw)
        ints.append(i); errs.append(e)
    out["unit/site_integral"] = numpy.array(ints)
    out["unit/site_error"] = numpy.array(errs)
    out["unit/sum_integral"] = numpy.array([sum(i) for i in ints])
    # --- synthetic ragged loci (NaN, zeros, fast sites that force multi-panel QAGS)
    for tag, kw, T, epochs in [
        ("synA", dict(seed=11, n_loci=12, smin=1, smax=700, T=1000), 1000, EPOCHS),
        ("synB", dict(seed=12, n_loci=6, smin=50, smax=300, T=2500, fast_frac=0.2), 2500,
         [[0, 10], [5, 60], [20, 100], [164, 174], [0, 2500]]),
    ]:
        loci = synthetic_loci(**kw)
        loci[0] = loci[0].copy(); loci[0][0] = numpy.inf           # +Inf site: dropped everywhere
        tv = ref.get_time(0, T)
        times = [0, 1, T // 3, T - 1, -1]
        offsets = numpy.zeros(len(loci) + 1, dtype=numpy.int64)
        offsets[1:] = numpy.cumsum([len(x) for x in loci])
        out[tag + "/offsets"] = offsets
        out[tag + "/rates"] = numpy.concatenate(loci)
        out[tag + "/T"] = numpy.int64(T)
        out[tag + "/times"] = numpy.array(times, dtype=numpy.int64)
        out[tag + "/epochs"] = numpy.array(epochs, dtype=numpy.int64)
        nets, pts, igs, ers, means = [], [], [], [], []
        for lam in loci:
            mean_rate, pi, pi_net, pi_times, pi_epochs = ref_worker(ref, tv, lam, times, epochs)
            nets.append(pi_net)
            pts.append([float(numpy.nansum(pi, axis=1)[t]) for t in times])
            i, e = epochs_to_arrays(pi_epochs, epochs)
            igs.append(i); ers.append(e); means.append(float(mean_rate))
        out[tag + "/pi_net"] = numpy.array(nets)
        out[tag + "/pi_times"] = numpy.array(pts)
        out[tag + "/integral"] = numpy.array(igs)
        out[tag + "/error"] = numpy.array(ers)
        out[tag + "/mean_rate"] = numpy.array(means)
    numpy.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    # --- the reference's data fixtures the host-side tests read (small text files, copied verbatim)
    import shutil
    os.makedirs(os.path.join(HERE, "ref-data"), exist_ok=True)
    for name in ["test-uniform-draw-weights.rates.json", "chr1_918.nex", "informativeness_cutoff.nex", "Euteleost.tree"]:
        shutil.copyfile(os.path.join(TD, name), os.path.join(HERE, "ref-data", name))
    # --- the reference's sqlite sink run on the config-1 worker result: DDL + every row it writes
    #     (tapir/db.py is imported unmodified; its py2 dict.iteritems is supplied by a dict subclass)
    import sqlite3
    import tempfile
    spec = importlib.util.spec_from_file_location("ref_db", os.path.join(REF, "tapir", "db.py"))
    ref_db = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_db)

    tmp = os.path.join(tempfile.mkdtemp(), "ref.sqlite")
    conn, c = ref_db.create_probe_db(tmp)
    ref_db.insert_pi_data(conn, c, [("/some/dir/locus-one.rates",) + cfg1])
    conn.commit()
    db_golden = {
        "ddl": [r[0] for r in conn.execute("SELECT sql FROM sqlite_master WHERE type='table' ORDER BY name")],
        "loci": conn.execute("SELECT * FROM loci").fetchall(),
        "net": conn.execute("SELECT * FROM net ORDER BY time").fetchall(),
        "discrete": conn.execute("SELECT * FROM discrete ORDER BY time").fetchall(),
        "interval": conn.execute("SELECT * FROM interval ORDER BY interval").fetchall(),
    }
    conn.close()
    with open(os.path.join(HERE, "reference_db_golden.json"), "w") as fh:
        json.dump(db_golden, fh, indent=1)
    meta = {
        "numpy": numpy.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
        "epochs": EPOCHS, "times": TIMES,
        "tree": open(os.path.join(TD, "Euteleost.tree")).read().strip(),
        "tree_expected_scaled": "(danRer6:1.74,(oryLat2:1,(gasAcu1:0.93,(fr2:0.37,tetNig2:0.37):0.56):0.07):0.74);",
        "phydesign_net_t10_20_50": [0.03448, 0.01111, 0.02293],
        "phydesign_integrals": [0.93453, 0.10628, 0.05855, 0.12638, 1.03698, 2.08840],
        "source": "generated by tests/golden/make_golden.py from /root/reference/tapir/compute.py (unmodified, shim import)",
    }
    with open(os.path.join(HERE, "reference_golden.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    main()
