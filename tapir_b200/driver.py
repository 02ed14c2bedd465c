"""Batched multi-locus driver: tapir_compute.py's main() and worker() (bin/tapir_compute.py:84-177) with
the per-locus numpy/scipy body replaced by ONE device pass over all loci.

The command line, the glob rules, the output directory rule, the Tree_*.newick side file, the rewritten
*.rates JSON and the sqlite database are the reference's.  ``--multiprocessing`` is accepted and means
"parse the rate files on all host cores and spread the loci over every visible GPU" (one host thread and one
context per device); without it one GPU is used.  The arithmetic always runs on the GPU.
"""
import argparse
import multiprocessing
import os
import sys
import time
from subprocess import PIPE, Popen

import numpy

from . import base, compute, db
from .engine import device_count

# wall-clock split of the last main(): loci, skipped, ingest_s, gpu_s, sqlite_s (tools/cli_walltime.py)
LAST_TIMING = {}


def get_args(argv=None):
    """Same options as bin/tapir_compute.py:18-53."""
    parser = argparse.ArgumentParser(description="""tapir:  compute the
            phylogenetic informativeness of DNA loci""")
    parser.add_argument('alignments', help="The folder of alignments", action=base.FullPaths, type=base.is_dir)
    parser.add_argument('tree', help="The input tree", action=base.FullPaths)
    required = parser.add_argument_group("required arguments")
    required.add_argument('--times', help="""Comma-separated list of start
        times of interest (MYA)""", type=base.get_list_from_ints, required=True)
    required.add_argument('--intervals', help="""Comma-separated list of
        interval ranges of interest (i,e. in MYA)""", type=base.get_list_from_ranges, required=True)
    parser.add_argument('--tree-format', help="The format of the tree", dest='tree_format',
                        choices=['nexus', 'newick'], default='newick')
    parser.add_argument('--output', help="The path to the output directory", default=os.getcwd(),
                        action=base.FullPaths, type=base.is_dir)
    parser.add_argument('--hyphy', help="The path to hyphy (if not in $PATH)", default="hyphy2")
    parser.add_argument('--template', help="""Path to the hypy
        temlate file if in non-standard location""", default=None)
    parser.add_argument('--threshold', help="""Minimum number of taxa without
        a gap for a site to be considered informative""", default=3, type=int)
    parser.add_argument('--multiprocessing', help="""Enable parallel
        calculation of rates""", default=False, action='store_true')
    parser.add_argument('--site-rates', default=False, action='store_true',
                        help="Use previously calculated site rates")
    parser.add_argument('--subset-pi-map-file', help="""Calculate PI for a
        subset of sites. If specified, this should be a tab-delimited file
        with the alignment file name in the 1st column, the start of the
        interval (0-offset) in the 2nd column, and the end of the interval in
        the 3rd column.""")
    # extension (not in the reference): read the rate files with the library's multi-threaded C++ reader and
    # leave them untouched, instead of json.load + rewrite-with-corrected_rates per file (compute.py:30-43)
    parser.add_argument('--no-rate-rewrite', default=False, action='store_true',
                        help="With --site-rates: do not rewrite the *.rates files with corrected rates (fast batched ingest)")
    # extension: the `net` table holds one row per locus and grid point (10^9 rows for 500,000 loci on a 2,048-point
    # grid); the plotting scripts that read it (tapir_plot.py, tapir_matplot.py) are optional consumers
    parser.add_argument('--no-net-table', default=False, action='store_true',
                        help="Do not fill the per-time-point `net` table (loci, discrete and interval are written as usual)")
    return parser.parse_args(argv)


def welcome_message():
    lines = ["tapir: high-throughput estimates of", "phylogenetic informativeness", "",
             "(c) 2011 Brant Faircloth, Jonathan Chang,", "Mike Alfaro", "",
             "PhyDesign was created by the Townsend Lab", "(http://phydesign.townsend.yale.edu)", "",
             "When using tapir, please cite:", "",
             "  - J.P. Townsend, 2007. Profiling", "    phylogenetic informativeness. Systematic",
             "    Biology, 56(2), 222-231.", "",
             "  - Pond, S.L.K., Frost, S.D.W., and S.V. Muse,", "    2005. Hyphy: hypothesis testing using",
             "    phylogenies. Bioinformatics, 21(5), 676-9.", "",
             "PI arithmetic: tapir_b200 (libtapir_cuda, sm_100a)"]
    bar = "    " + "*" * 51
    body = "\n".join("    * {0:<47s} *".format(x) for x in [""] + lines + [""])
    return "\n" + bar + "\n" + body + "\n" + bar + "\n\n"


def load_rates(params):
    """The host half of worker() (bin/tapir_compute.py:90-108): run HyPhy when asked to, read the rate
    JSON, cull uninformative sites, apply the subset slice.  Returns (alignment, rates)."""
    (_time_vector, hyphy, template, towrite, output, correction, alignment, _times, _epochs, threshold, subsets) = params
    sys.stdout.write(".")
    sys.stdout.flush()
    if towrite:
        try:
            proc = Popen([hyphy, template], stdin=PIPE, stdout=PIPE)
        except OSError as err:
            if err.errno == 2:
                raise Exception("hyphy not found")
            raise Exception("couldn't communicate with hyphy: {0}".format(err))
        stdout, _ = proc.communicate(towrite.encode())
        if stdout.decode(errors="replace").startswith("Error"):
            raise Exception("hyphy error: {0}".format(stdout))
        rates = compute.parse_site_rates(output, correction=correction)
        good_sites = compute.get_informative_sites(alignment, threshold)
        rates = compute.cull_uninformative_rates(rates, good_sites)
    else:
        rates = compute.parse_site_rates(output, correction=correction)
    alignment_basename = os.path.basename(alignment)
    if alignment_basename in subsets:
        rates = rates[subsets[alignment_basename][0]:subsets[alignment_basename][1]]
    return alignment, rates


def worker(params):
    """Single-locus form with the reference's signature and 7-tuple result (bin/tapir_compute.py:84-123)."""
    alignment, rates = load_rates(params)
    time_vector, times, epochs = params[0], params[7], params[8]
    return compute.compute_loci([rates], time_vector, times, epochs, names=[alignment], keep_matrix=True)[0]


def load_batch(params, threads=1, native_ingest=False):
    """Host half of every worker() of a run -> (names, offsets, rates): the loci's rate vectors as one CSR pack.

    native_ingest (only for --site-rates parameters): the rate files are parsed by tapir_ingest_rates on `threads`
    C++ threads straight into the pack and are not rewritten; otherwise load_rates() per file (optionally on a thread
    pool), which keeps the reference's rewrite of every *.rates file."""
    from .engine import pack_csr
    names = [p[6] for p in params]
    subsets = params[0][10] if params else {}
    if native_ingest and all(p[3] is None for p in params):
        from . import ingest
        sys.stdout.write("." * len(params))
        sys.stdout.flush()
        offsets, rates = ingest.read_rate_files([p[4] for p in params], correction=params[0][5], threads=threads,
                                                pinned=not subsets)
        if not subsets:
            return names, offsets, rates
        vectors = []
        for l, name in enumerate(names):
            v = rates[offsets[l]:offsets[l + 1]]
            base_name = os.path.basename(name)
            if base_name in subsets:
                v = v[subsets[base_name][0]:subsets[base_name][1]]
            vectors.append(v)
    else:
        if threads > 1 and len(params) >= 64 and "fork" in multiprocessing.get_all_start_methods():
            # worker processes, as the reference's --multiprocessing (bin/tapir_compute.py:155-159): the JSON work is
            # pure Python, so threads only queue on the interpreter lock (measured: 8 threads 0.6x, 8 processes 4x one
            # core).  The children parse, rewrite and return the vectors; they never touch the GPU.
            with multiprocessing.get_context("fork").Pool(threads) as pool:
                loaded = pool.map(load_rates, params, chunksize=max(1, min(64, len(params) // (4 * threads))))
        elif threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=threads) as pool:
                loaded = list(pool.map(load_rates, params))
        else:
            loaded = [load_rates(p) for p in params]
        names = [a for a, _ in loaded]
        vectors = [r for _, r in loaded]
    offsets, rates = pack_csr(vectors, pinned=True)
    return names, offsets, rates


def run_batch_arrays(params, threads=1, native_ingest=False, devices=None, want_profile=True):
    """All loci of a run: host-side ingest, then ONE device pass -> (names, BatchResult, skipped).

    A locus without any finite site rate (the reference crashes on it inside numpy.vectorize and takes the whole
    run down) is left out of the pass and returned in `skipped`; names / result rows cover the remaining loci in
    glob order."""
    if not params:
        return [], None, []
    t0 = time.perf_counter()
    names, offsets, rates = load_batch(params, threads=threads, native_ingest=native_ingest)
    LAST_TIMING["ingest_s"] = time.perf_counter() - t0
    bad = compute.loci_without_finite_rate(offsets, rates)
    skipped = [names[int(l)] for l in bad]
    if bad.size:
        from .engine import shard_csr
        keep = numpy.setdiff1d(numpy.arange(len(names)), bad)
        offsets, rates = shard_csr(offsets, rates, keep, pinned=True)
        names = [names[int(l)] for l in keep]
    time_vector, times, epochs = params[0][0], params[0][7], params[0][8]
    t0 = time.perf_counter()
    res = compute.compute_batch(offsets, rates, time_vector, times, epochs, devices=devices, want_profile=want_profile)
    LAST_TIMING["gpu_s"] = time.perf_counter() - t0
    return names, res, skipped


def run_batch(params, threads=1, native_ingest=False, devices=None):
    """run_batch_arrays in the reference's result format: one worker() 7-tuple per locus (what map(worker, params)
    returns, bin/tapir_compute.py:160), loci without a finite rate left out."""
    names, res, _ = run_batch_arrays(params, threads=threads, native_ingest=native_ingest, devices=devices)
    if res is None:
        return []
    times, epochs = params[0][7], params[0][8]
    labels = ["{0}-{1}".format(s[0], s[1]) for s in epochs]
    pi_times, integral, abserr = res.pi_times.tolist(), res.integral.tolist(), res.abserr.tolist()
    mean_rate, pi_net = res.mean_rate.tolist(), numpy.array(res.pi_net)
    out = []
    for l, name in enumerate(names):
        pi_epochs = {lab: {"sum(integral)": integral[l][k], "sum(error)": abserr[l][k]} for k, lab in enumerate(labels)}
        out.append((name, None, mean_rate[l], None, pi_net[l], dict(zip(times, pi_times[l])), pi_epochs))
    return out


def get_hyphy_conf():
    """Path of the HyPhy batch template; the template itself ships with the reference package
    (tapir/data/models_and_rates.bf) and is out of scope here: point --template (or $TAPIR_HYPHY_TEMPLATE) at it."""
    return os.environ.get("TAPIR_HYPHY_TEMPLATE", "models_and_rates.bf")


def main(argv=None):
    args = get_args(argv)
    print(welcome_message())
    args.output = base.create_unique_dir(args.output)
    tree_depth, correction, tree = compute.correct_branch_lengths(args.tree, args.tree_format, d=args.output)
    time_vector = compute.get_time(0, int(tree_depth))
    subset_pi = dict()
    if args.subset_pi_map_file:
        subset_pi = dict(base.parse_subset_map_file(args.subset_pi_map_file))
    params = []
    template = args.template if args.template else get_hyphy_conf()
    if not args.site_rates:
        print("\nEstimating site rates and PI for files:")
        for alignment in base.get_files(args.alignments, '*.nex,*.nexus'):
            output = os.path.join(args.output, os.path.basename(alignment) + '.rates')
            towrite = "\n".join([alignment, tree, output])
            params.append([time_vector, args.hyphy, template, towrite, output, correction, alignment,
                           args.times, args.intervals, args.threshold, subset_pi])
    else:
        print("Estimating PI for files (--site-rate option):")
        for rate_file in base.get_files(args.alignments, '*.rates'):
            params.append([time_vector, args.hyphy, template, None, rate_file, correction, rate_file,
                           args.times, args.intervals, args.threshold, subset_pi])
    # --multiprocessing: all host cores for the ingest (as the reference's Pool) and every visible GPU for the math
    threads = max(1, (os.cpu_count() or 2) - 1) if args.multiprocessing else 1
    devices = list(range(max(1, device_count()))) if args.multiprocessing else None
    t0 = time.perf_counter()
    names, res, skipped = run_batch_arrays(params, threads=threads, native_ingest=args.site_rates and args.no_rate_rewrite,
                                           devices=devices, want_profile=not args.no_net_table)
    if skipped:
        sys.stdout.write("\nSkipped {0} loci without any finite site rate: {1}{2}\n".format(
            len(skipped), ", ".join(db.locus_name(s) for s in skipped[:10]), " ..." if len(skipped) > 10 else ""))
    t2 = time.perf_counter()
    db_name = os.path.join(args.output, 'phylogenetic-informativeness.sqlite')
    sys.stdout.write("\nStoring results in {0}...".format(db_name))
    sys.stdout.flush()
    conn, c = db.create_probe_db(db_name)
    # the bulk writer: same rows, same order as insert_pi_data over worker() tuples (tapir/db.py:44-61).  With
    # --multiprocessing a large `net` table goes through worker processes; the foreign keys (ids generated by this very
    # call) are not re-checked row by row while it is filled, which lets SQLite append the shards table to table.
    if threads > 1:
        c.execute("PRAGMA foreign_keys = OFF")
    db.insert_batch(conn, c, names, res, args.times, args.intervals, with_net=not args.no_net_table,
                    workers=threads, tmp_dir=args.output)
    conn.commit()
    if threads > 1:
        db.detach_shards(conn)
        c.execute("PRAGMA foreign_keys = ON")
    t3 = time.perf_counter()
    LAST_TIMING.update(loci=len(names), skipped=len(skipped), ingest_and_gpu_s=t2 - t0, sqlite_s=t3 - t2)
    sys.stdout.write("DONE")
    sys.stdout.flush()
    print("\n")
    c.close()
    conn.close()
    return 0
