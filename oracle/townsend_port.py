"""numpy/scipy port of the reference's per-locus hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to /root/reference).
The arithmetic deliberately goes through the same third-party calls as the reference
(numpy ufuncs, numpy.nansum, scipy.integrate.quad, python builtin sum) so that results are
the reference's results on this interpreter; tests/golden/ holds vectors generated from the
reference's own compute.py to prove it.
"""
import json
import os

import numpy as np
from scipy import integrate

QUAD_EPS = 1.49e-8  # scipy.integrate.quad defaults used by compute.py:52
QUAD_LIMIT = 50


def site_rates_from_json(path, correction=1):
    """tapir/compute.py:24-44 (read side only): data['sites']['rates'][i]['rate'] / correction."""
    with open(path) as fh:
        doc = json.load(fh)
    raw = np.array([rec["rate"] for rec in doc["sites"]["rates"]])
    return raw / correction


def pi_matrix(time_column, rates):
    """tapir/compute.py:46-48.  time_column (T,1) int64, rates (S,) -> (T,S) float64."""
    return 16 * (rates ** 2) * time_column * np.exp(-(4 * rates * time_column))


def time_column(start, stop, step=1):
    """tapir/compute.py:54-57."""
    return np.reshape(np.array(range(start, stop, step)), (-1, 1))


def site_integral(start, stop, rate):
    """tapir/compute.py:50-52: QUADPACK QAGS via scipy, (integral, abserr)."""
    return integrate.quad(pi_matrix, start, stop, args=(rate))


def epoch_integrals(finite_rates, epochs):
    """tapir/compute.py:81-94: per epoch, python-sum of per-site quad() results and errors."""
    per_site = np.vectorize(site_integral)
    out = {}
    for span in epochs:
        label = "{0}-{1}".format(span[0], span[1])
        assert span[0] < span[1], "Start time [{0}] is sooner than end time [{1}]".format(span[0], span[1])
        integral, error = per_site(span[0], span[1], finite_rates)
        out[label] = {"sum(integral)": sum(integral), "sum(error)": sum(error)}
    return out


def picks(pi, times):
    """tapir/compute.py:76-79: nansum over sites, fancy-indexed by grid index."""
    return dict(zip(times, np.nansum(pi, axis=1)[times]))


def locus_worker(time_col, rates, times, epochs, keep_matrix=True):
    """bin/tapir_compute.py:104-123 for an already-parsed rate vector.

    Returns (rates, mean_rate, pi, pi_net, pi_times, pi_epochs) -- the reference's 7-tuple
    minus the alignment name.
    """
    mean_rate = np.mean(np.ma.masked_array(rates, np.isnan(rates)))
    pi = pi_matrix(time_col, rates)
    pi_net = np.nansum(pi, axis=1)
    pi_times = picks(pi, times)
    pi_epochs = epoch_integrals(rates[np.isfinite(rates)], epochs)
    return rates, mean_rate, (pi if keep_matrix else None), pi_net, pi_times, pi_epochs


def argmax_epoch(pi_epochs, epochs):
    """README.rst:178-185: the interval whose summed integral is the per-locus maximum
    (first in --intervals order on ties)."""
    labels = ["{0}-{1}".format(a, b) for a, b in epochs]
    vals = [pi_epochs[k]["sum(integral)"] for k in labels]
    return int(np.argmax(vals))


def tree_depth_and_correction(newick):
    """tapir/compute.py:59-68 on a plain newick string (dendropy semantics restated):
    depth = max root-to-tip path, mean_bl = total edge length / (2*n_leaves - 3),
    cf = 10**len(str(int(mean_bl + 0.5))) if that length > 1 else 1.  Small trees only."""
    s = newick.strip()
    stats = {"total": 0.0, "leaves": 0}

    def label_end(k):
        while k < len(s) and s[k] not in ",();":
            k += 1
        return k

    def edge(tok):
        bl = float(tok.split(":")[1]) if ":" in tok else 0.0
        stats["total"] += bl
        return bl

    def node(k):
        if s[k] != "(":
            j = label_end(k)
            stats["leaves"] += 1
            return edge(s[k:j]), j
        k += 1
        deepest = 0.0
        while True:
            d, k = node(k)
            deepest = max(deepest, d)
            if s[k] == ",":
                k += 1
            elif s[k] == ")":
                k += 1
                break
        j = label_end(k)
        return deepest + edge(s[k:j]), j

    depth, _ = node(0)
    mean_bl = stats["total"] / (2 * stats["leaves"] - 3)
    width = len(str(int(mean_bl + 0.5)))
    return depth, (10 ** width if width > 1 else 1)
