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


_INF = float("inf")
try:
    from json.encoder import c_encode_basestring_ascii as _json_str
except ImportError:                                   # pragma: no cover
    from json.encoder import py_encode_basestring_ascii as _json_str


def _json_scalar(o):
    """Text of a JSON scalar as json.dumps writes it, or None when `o` is a container / needs json's own handling."""
    if isinstance(o, str):
        return _json_str(o)
    if o is None:
        return "null"
    if o is True:
        return "true"
    if o is False:
        return "false"
    if isinstance(o, int):
        return int.__repr__(o)
    if isinstance(o, float):
        if o != o:
            return "NaN"
        if o == _INF:
            return "Infinity"
        if o == -_INF:
            return "-Infinity"
        return float.__repr__(o)
    return None


def dumps_indent4(obj):
    """``json.dumps(obj, indent=4)``, byte for byte, for the documents the rate files hold (dicts, lists, strings,
    numbers, booleans, None), several times faster: with an indent the standard library falls back to its
    pure-Python generator encoder, which dominates the rewrite of a rate file (5 ms of 5.5 ms for 500 sites).
    Anything this writer does not know (other key or value types) goes to json.dumps itself."""
    out = []
    pads = ["\n"]

    def pad(level):
        while len(pads) <= level:
            pads.append("\n" + " " * (4 * len(pads)))
        return pads[level]

    def column(values):
        """Texts of one column of a table (all values of one plain scalar type), or None."""
        kinds = set(map(type, values))
        if len(kinds) != 1:
            return None
        kind = kinds.pop()
        if kind is float:
            texts = list(map(float.__repr__, values))
            for i, t in enumerate(texts):
                if t[-1] in "nf":                           # nan, inf, -inf
                    texts[i] = _json_scalar(values[i])
            return texts
        if kind is int:
            return list(map(int.__repr__, values))
        if kind is str:
            return list(map(_json_str, values))
        if kind is bool or values[0] is None:
            return list(map(_json_scalar, values))
        return None

    def table(rows, level):
        """A list of dicts that all have the same string keys and plain scalar values (the `rates` and
        `corrected_rates` lists of a rate file): one format string per row, the columns converted by type."""
        first = rows[0]
        if type(first) is not dict or not first:
            return False
        keys = tuple(first)
        for row in rows:
            if type(row) is not dict or tuple(row) != keys:
                return False
        if not all(type(k) is str for k in keys):
            return False
        cols = []
        for k in keys:
            texts = column([row[k] for row in rows])
            if texts is None:
                return False
            cols.append(texts)
        p1, p2 = pad(level + 1), pad(level + 2)
        fmt = "{" + ",".join(p2 + _json_str(k).replace("%", "%%") + ": %s" for k in keys) + p1 + "}"
        out.append("[" + p1 + ("," + p1).join(fmt % t for t in zip(*cols)) + pad(level) + "]")
        return True

    def emit(o, level):
        text = _json_scalar(o)
        if text is not None:
            out.append(text)
        elif isinstance(o, (list, tuple)):
            if not o:
                out.append("[]")
                return
            if len(o) >= 8 and table(o, level):
                return
            inner = pad(level + 1)
            sep = "[" + inner
            for item in o:
                out.append(sep)
                emit(item, level + 1)
                sep = "," + inner
            out.append(pad(level) + "]")
        elif isinstance(o, dict):
            if not o:
                out.append("{}")
                return
            inner = pad(level + 1)
            sep = "{" + inner
            for key, value in o.items():
                if not isinstance(key, str):
                    raise TypeError("non-string key")      # json's key coercions: leave them to json
                out.append(sep)
                out.append(_json_str(key))
                out.append(": ")
                emit(value, level + 1)
                sep = "," + inner
            out.append(pad(level) + "}")
        else:
            raise TypeError("not a plain JSON value")

    try:
        emit(obj, 0)
    except (TypeError, RecursionError):
        return json.dumps(obj, indent=4)
    return "".join(out)


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
            handle.write(dumps_indent4(data))                 # = json.dump(data, handle, indent=4), byte for byte
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
    if not numpy.isfinite(rates).all():
        # the reference integrates whatever it is given: a NaN rate makes sum(integral) NaN, an infinite one makes quad
        # fail.  worker() always filters first (bin/tapir_compute.py:122); the batched engine treats non-finite rates
        # as culled sites, so an unfiltered vector is refused here rather than silently summed without them.
        raise ValueError("get_net_integral_for_epochs expects finite rates: filter with rates[numpy.isfinite(rates)] "
                         "as worker() does (bin/tapir_compute.py:122)")
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


def compute_batch(offsets, rates, time_vector, times, epochs, engine=None, devices=None, want_profile=True):
    """The body of worker() (bin/tapir_compute.py:104-123) for MANY loci in one device pass, array in / array out.

    (offsets, rates): CSR pack of the loci's corrected rates (NaN = culled site).  Returns an engine.BatchResult
    whose row l is locus l: pi_net (T values, the `net` table), pi_times (`discrete`), integral / abserr
    (`interval`), mean_rate, ... -- no per-locus Python object is built.  `devices`: list of CUDA device indices to
    spread the loci over (one host thread and one context per device); default: the process-wide engine."""
    time_vector = numpy.asarray(time_vector)
    T = int(time_vector.reshape(-1).size)
    if T and not numpy.array_equal(time_vector.reshape(-1), numpy.arange(T)):
        raise ValueError("the batched path expects the grid get_time(0, T) (bin/tapir_compute.py:135)")
    for span in epochs:
        assert span[0] < span[1], "Start time [{0}] is sooner than end time [{1}]".format(span[0], span[1])
    if times:
        numpy.empty(T)[list(times)]                     # numpy's IndexError for an out-of-range time (compute.py:78)
    span_list = [[s[0], s[1]] for s in epochs]
    if devices is not None and len(devices) > 1:
        from . import sharded
        return sharded.run_threaded(offsets, rates, T, list(times), span_list, devices, want_profile=want_profile)
    eng = engine or get_engine()
    return eng.run(offsets, rates, T, list(times), span_list, want_profile=want_profile)


def loci_without_finite_rate(offsets, rates):
    """Indices of the loci that have no finite site rate at all (empty, or every site culled)."""
    offsets = numpy.asarray(offsets, dtype=numpy.int64)
    fin = numpy.concatenate([[0], numpy.cumsum(numpy.isfinite(rates), dtype=numpy.int64)])
    return numpy.nonzero(fin[offsets[1:]] == fin[offsets[:-1]])[0]


def compute_loci(rate_vectors, time_vector, times, epochs, names=None, keep_matrix=False, engine=None, devices=None):
    """compute_batch in the reference's own result format: a list of worker() 7-tuples
    (name, rates, mean_rate, pi, pi_net, pi_times, pi_epochs) in input order, exactly what insert_pi_data
    (tapir/db.py:44-61) consumes; `pi` (the (T,S) matrix the reference pickles back to the parent) is None unless
    keep_matrix is set.  A locus without any finite rate raises ValueError, as the reference dies on it inside
    numpy.vectorize; the batched driver (driver.run_batch) skips and reports such loci instead."""
    n = len(rate_vectors)
    names = list(names) if names is not None else [None] * n
    vectors = [numpy.ascontiguousarray(v, dtype=numpy.float64).reshape(-1) for v in rate_vectors]
    offsets, rates = _engine_mod.pack_csr(vectors, pinned=not (devices and len(devices) > 1))
    bad = loci_without_finite_rate(offsets, rates)
    if bad.size:
        raise ValueError("locus {0!r} has no finite site rate".format(names[int(bad[0])]))
    res = compute_batch(offsets, rates, time_vector, times, epochs, engine=engine, devices=devices)
    labels = ["{0}-{1}".format(s[0], s[1]) for s in epochs]
    tcol = numpy.asarray(time_vector).reshape(-1, 1)
    pi_times, integral, abserr = res.pi_times.tolist(), res.integral.tolist(), res.abserr.tolist()
    mean_rate = res.mean_rate.tolist()
    pi_net = numpy.array(res.pi_net)                      # one copy out of the (possibly pinned) result buffer
    times = list(times)
    out = []
    for l in range(n):
        pi = get_townsend_pi(tcol, vectors[l]) if keep_matrix else None
        pi_epochs = {lab: {"sum(integral)": integral[l][k], "sum(error)": abserr[l][k]} for k, lab in enumerate(labels)}
        out.append((names[l], vectors[l], mean_rate[l], pi, pi_net[l], dict(zip(times, pi_times[l])), pi_epochs))
    return out
