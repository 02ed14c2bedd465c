"""Drop-in mirror of ``tapir.compute`` (tapir/compute.py:24-110) whose arithmetic runs in libtapir_cuda.

Every function keeps the reference's name, arguments, return type and error behaviour, so code (and
tests) written against ``tapir.compute`` can switch with ``import tapir_b200.compute as compute``.
The numpy/scipy math is replaced by calls through the C ABI; numpy is used for host arrays only.
There is no CPU fallback: without a CUDA device the first compute call raises ``TapirCudaError``.

Added for the batched driver (bin/tapir_compute.py:153-164 mapped over ALL loci in one call):
``compute_loci`` -- many rate vectors in, one worker()-style 7-tuple per locus out.
"""
import json
import os
import time

import numpy

from . import engine as _engine_mod
from . import tree as _tree

_ENGINE = None


def get_engine():
    """The process-wide Engine (device from $TAPIR_CUDA_DEVICE, default 0), created on first use."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = _engine_mod.Engine(int(os.environ.get("TAPIR_CUDA_DEVICE", "0")))
    return _ENGINE


def parse_site_rates(rate_file, correction=1, test=False, count=0):
    """Site-rate JSON written by the HyPhy template -> vector of rates / correction (compute.py:24-44).

    As the reference does, the file is rewritten with a ``corrected_rates`` list unless ``test`` is set.
    A file that cannot be opened is retried up to three times, 0.1 s apart, then IOError is raised
    (the reference's retry discards its own result; here the retry is effective)."""
    while True:
        try:
            with open(rate_file, "r") as handle:
                data = json.load(handle)
            break
        except IOError as err:
            if count <= 3:
                count += 1
                time.sleep(0.1)
                continue
            raise IOError("Cannot open {0}: {1}".format(rate_file, err))
    rates = numpy.array([line["rate"] for line in data["sites"]["rates"]])
    corrected = rates / correction
    if not test:
        data["sites"]["corrected_rates"] = [{"site": k + 1, "rate": v} for k, v in enumerate(corrected.tolist())]
        with open(rate_file, "w") as handle:
            json.dump(data, handle, indent=4)
    return corrected


def get_townsend_pi(time, rates):
    """Townsend et al. equation 10, 16 r^2 t exp(-4 r t), with numpy's broadcasting of a time column
    (T,1) against a rate vector (S,) -> (T,S) (compute.py:46-48).  Scalars are accepted on either side."""
    t = numpy.asarray(time, dtype=numpy.float64)
    r = numpy.asarray(rates, dtype=numpy.float64)
    shape = numpy.broadcast_shapes(t.shape, r.shape)
    column = t.ndim == 2 and t.shape[1] == 1 and r.ndim <= 1
    if not (column or t.ndim == 0 or r.ndim == 0):
        raise ValueError("get_townsend_pi: time must be a (T,1) column or a scalar when rates is an array")
    out = get_engine().pi_matrix(t.reshape(-1), r.reshape(-1))
    out = out.reshape(shape)
    return out if out.ndim else float(out)


def get_integral_over_times(start, stop, rate):
    """QUADPACK QAGS integral of the PI curve of one site over [start, stop] -> (integral, abserr), what
    scipy.integrate.quad returns for compute.py:50-52."""
    integ, err, _ = get_engine().site_integrals(start, stop, numpy.array([rate], dtype=numpy.float64))
    return float(integ[0]), float(err[0])


def get_time(start, stop, step=1):
    """Column of grid times (compute.py:54-57)."""
    return numpy.reshape(numpy.array(range(start, stop, step)), (-1, 1))


def correct_branch_lengths(tree_file, format, d=""):
    """Scale branch lengths by a power of ten chosen from the mean branch length and write
    Tree_{cf}_{depth}.newick into `d`; returns (depth, correction_factor, path) (compute.py:59-74).
    The depth is taken BEFORE scaling, as in the reference."""
    tree = _tree.read_tree(tree_file, format)
    depth = tree.depth()
    mean_branch_length = tree.length() / (2 * len(tree.leaf_nodes()) - 3)
    string_len = len(str(int(mean_branch_length + 0.5)))
    correction_factor = 10 ** string_len if string_len > 1 else 1
    tree.scale_edges(correction_factor)
    pth = os.path.join(d, "Tree_{0}_{1}.newick".format(correction_factor, depth))
    with open(pth, "w") as handle:
        handle.write(tree.to_newick() + "\n")
    return depth, correction_factor, pth


def get_net_pi_for_periods(pi, times):
    """nansum over sites of a (T,S) PI matrix, picked at the grid indices `times` -> {time: value}
    (compute.py:76-79; numpy fancy-index rules: negative indices wrap, out of range raises IndexError)."""
    sums = get_engine().nansum_rows(numpy.asarray(pi, dtype=numpy.float64))[times]
    return dict(zip(times, sums))


def get_net_integral_for_epochs(rates, epochs):
    """{'a-b': {'sum(integral)': .., 'sum(error)': ..}} over the given finite rates (compute.py:81-94)."""
    rates = numpy.ascontiguousarray(rates, dtype=numpy.float64).reshape(-1)
    for span in epochs:
        assert span[0] < span[1], "Start time [{0}] is sooner than end time [{1}]".format(span[0], span[1])
    if rates.size == 0:
        # numpy.vectorize on an empty vector raises in the reference
        raise ValueError("cannot integrate over an empty rate vector")
    epochs_results = {}
    if not epochs:
        return epochs_results
    offsets = numpy.array([0, rates.size], dtype=numpy.int64)
    res = get_engine().run(offsets, rates, 0, [], [[s[0], s[1]] for s in epochs], want_profile=False)
    for k, span in enumerate(epochs):
        name = "{0}-{1}".format(span[0], span[1])
        epochs_results[name] = {"sum(integral)": float(res.integral[0, k]), "sum(error)": float(res.abserr[0, k])}
    return epochs_results


def _read_nexus_matrix(path):
    """Sequences of the first DATA/CHARACTERS matrix of a nexus file (interleaved blocks are joined)."""
    with open(path) as handle:
        text = handle.read()
    import re
    text = re.sub(r"\[[^\]]*\]", "", text)
    m = re.search(r"\bmatrix\b(.*?);", text, flags=re.I | re.S)
    if not m:
        raise ValueError("no MATRIX in {0}".format(path))
    seqs, order = {}, []
    for line in m.group(1).splitlines():
        line = line.strip()
        if not line:
            continue
        if line[0] in "'\"":
            q = line[0]
            end = line.index(q, 1)
            name, rest = line[1:end], line[end + 1:]
        else:
            parts = line.split(None, 1)
            name, rest = parts[0], (parts[1] if len(parts) > 1 else "")
        if name not in seqs:
            seqs[name] = []
            order.append(name)
        seqs[name].append("".join(rest.split()))
    return [("".join(seqs[name])) for name in order]


def get_informative_sites(alignment, threshold=4):
    """1.0 where at least `threshold` taxa have an unambiguous base (A/C/G/T), NaN elsewhere
    (compute.py:96-106)."""
    rows = _read_nexus_matrix(alignment)
    if not rows:
        return numpy.array([])
    width = len(rows[0])
    counts = numpy.zeros(width, dtype=numpy.int64)
    for seq in rows:
        assert len(seq) == width                      # should all have equal lengths
        counts += numpy.isin(numpy.frombuffer(seq.upper().encode("ascii"), dtype=numpy.uint8),
                             numpy.frombuffer(b"ATGC", dtype=numpy.uint8))
    return numpy.where(counts >= threshold, 1.0, numpy.nan)


def cull_uninformative_rates(rates, inform):
    """NaN-out the rates of uninformative sites (compute.py:108-110)."""
    return rates * inform


def compute_loci(rate_vectors, time_vector, times, epochs, names=None, keep_matrix=False, engine=None, devices=None):
    """The body of worker() (bin/tapir_compute.py:104-123) for MANY loci in one device pass.

    rate_vectors: sequence of 1-D rate arrays (NaN = culled site), already corrected/subset.
    Returns a list of 7-tuples (name, rates, mean_rate, pi, pi_net, pi_times, pi_epochs) in input order,
    exactly what insert_pi_data (tapir/db.py:44-61) consumes; `pi` (the (T,S) matrix the reference
    pickles back to the parent) is None unless keep_matrix is set.  `devices`: list of CUDA device indices to
    spread the loci over (one host thread and one context per device); default: the process-wide engine."""
    eng = engine or (get_engine() if not (devices and len(devices) > 1) else None)
    n = len(rate_vectors)
    names = list(names) if names is not None else [None] * n
    time_vector = numpy.asarray(time_vector)
    T = int(time_vector.reshape(-1).size)
    if T and not numpy.array_equal(time_vector.reshape(-1), numpy.arange(T)):
        raise ValueError("the batched path expects the grid get_time(0, T) (bin/tapir_compute.py:135)")
    for span in epochs:
        assert span[0] < span[1], "Start time [{0}] is sooner than end time [{1}]".format(span[0], span[1])
    vectors = [numpy.ascontiguousarray(v, dtype=numpy.float64).reshape(-1) for v in rate_vectors]
    for name, v in zip(names, vectors):
        if not numpy.isfinite(v).any():
            # the reference dies inside numpy.vectorize on an empty finite-rate vector
            raise ValueError("locus {0!r} has no finite site rate".format(name))
    if times:
        numpy.empty(T)[list(times)]                     # numpy's IndexError for an out-of-range time (compute.py:78)
    span_list = [[s[0], s[1]] for s in epochs]
    if devices is not None and len(devices) > 1:
        from . import sharded
        offsets, rates = _engine_mod.pack_csr(vectors, pinned=False)
        res = sharded.run_threaded(offsets, rates, T, list(times), span_list, devices)
    else:
        offsets, rates = _engine_mod.pack_csr(vectors, pinned=True)
        res = eng.run(offsets, rates, T, list(times), span_list, want_profile=True)
    labels = ["{0}-{1}".format(s[0], s[1]) for s in epochs]
    out = []
    tcol = time_vector.reshape(-1, 1)
    for l in range(n):
        pi = get_townsend_pi(tcol, vectors[l]) if keep_matrix else None
        pi_times = dict(zip(times, res.pi_times[l].tolist()))
        pi_epochs = {lab: {"sum(integral)": float(res.integral[l, k]), "sum(error)": float(res.abserr[l, k])}
                     for k, lab in enumerate(labels)}
        out.append((names[l], vectors[l], float(res.mean_rate[l]), pi, numpy.array(res.pi_net[l]), pi_times, pi_epochs))
    return out
