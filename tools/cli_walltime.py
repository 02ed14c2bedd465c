#!/usr/bin/env python
"""Wall time of the command line on many rate files (GPU box): ingest / GPU / sqlite split of tapir_compute.py --site-rates
for N synthetic *.rates files (500 sites each, tree depth 1000 -> 1000-point grid), in three forms:
  reference-compatible   json.load + rewrite of every file, `net` table written (also with --multiprocessing)
  --no-rate-rewrite --multiprocessing          batched C++ ingest, every visible GPU
  ... --no-net-table                            without the T-rows-per-locus table
usage: python tools/cli_walltime.py [n_files=50000] [n_sites=500] > profiles/r02_cli_walltime.txt"""
import json
import multiprocessing as mp
import os
import shutil
import sqlite3
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 500
TREE = "(a:1000.0,(b:600.0,c:600.0):400.0);"     # depth 1000, mean branch 866.7 -> correction factor 1000


def write_files(job):
    d, l0, l1 = job
    for l in range(l0, l1):
        rng = np.random.default_rng([77, l])
        lam = rng.gamma(0.5, 0.004, S)
        lam[rng.random(S) < 0.3] = 0.0
        recs = ", ".join('{"site": %d, "subst": 0.0, "rate": %.4f, "ll": 0.0}' % (k + 1, v * 1000.0) for k, v in enumerate(lam))
        with open(os.path.join(d, "locus-%06d.rates" % l), "w") as fh:
            fh.write('{"sites": {"rates": [' + recs + ']}}')


def main():
    work = tempfile.mkdtemp(prefix="tapir_cli_")
    try:
        src = os.path.join(work, "rates")
        os.mkdir(src)
        tree = os.path.join(work, "in.tree")
        open(tree, "w").write(TREE + "\n")
        t0 = time.perf_counter()
        step = 500
        with mp.get_context("fork").Pool(max(1, (os.cpu_count() or 2) - 1)) as pool:
            pool.map(write_files, [(src, l, min(N, l + step)) for l in range(0, N, step)])
        print("# %d rate files x %d sites written in %.1f s (%d CPUs); tree depth 1000 -> T = 1000, 4 epochs, 4 times" % (
            N, S, time.perf_counter() - t0, os.cpu_count()))
        forms = [("reference-compatible (json.load + rewrite per file, net table)", [], min(N, 5000)),
                 ("reference-compatible --multiprocessing (worker processes)", ["--multiprocessing"], min(N, 5000)),
                 ("--no-rate-rewrite --multiprocessing", ["--no-rate-rewrite", "--multiprocessing"], N),
                 ("--no-rate-rewrite --multiprocessing --no-net-table", ["--no-rate-rewrite", "--multiprocessing", "--no-net-table"], N)]
        for label, extra, n_use in forms:
            d = src
            if n_use < N:                         # the slow form runs on a prefix of the files (copied: it rewrites them)
                d = os.path.join(work, "subset-%d" % len(os.listdir(work)))
                os.mkdir(d)
                for l in range(n_use):
                    shutil.copy(os.path.join(src, "locus-%06d.rates" % l), d)
            out = os.path.join(work, "out-%d" % len(os.listdir(work)))
            os.mkdir(out)
            argv = [d, tree, "--output", out, "--times=37,93,100,170", "--intervals=32-42,88-98,95-105,164-174", "--site-rates"] + extra
            # every form in a process of its own, as a user's command is: the wall time includes the CUDA context, and
            # the worker processes of --multiprocessing are forked before any context exists
            code = ("import json, sys, time; sys.path.insert(0, %r); from tapir_b200 import driver; "
                    "out = sys.stdout; sys.stdout = open('/dev/null', 'w'); t0 = time.perf_counter(); driver.main(%r); "
                    "w = time.perf_counter() - t0; sys.stdout = out; print(json.dumps([w, driver.LAST_TIMING]))" % (ROOT, argv))
            proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
            if proc.returncode != 0:
                print("%-70s failed: %s" % (label, proc.stderr.strip().split("\n")[-1][:200]))
                continue
            wall, tm = json.loads(proc.stdout.strip().split("\n")[-1])
            db = os.path.join(out, "phylogenetic-informativeness.sqlite")
            conn = sqlite3.connect(db)
            rows = {t: conn.execute("SELECT COUNT(*) FROM %s" % t).fetchone()[0] for t in ("loci", "net", "discrete", "interval")}
            conn.close()
            print("%-70s %6d loci: wall %7.2f s = ingest %6.2f + gpu %5.2f + sqlite %6.2f (+ tree/args %.2f)  -> %.0f loci/s   rows %s  db %.0f MB" % (
                label, tm["loci"], wall, tm["ingest_s"], tm["gpu_s"], tm["sqlite_s"],
                wall - tm["ingest_s"] - tm["gpu_s"] - tm["sqlite_s"], tm["loci"] / wall, rows, os.path.getsize(db) / 1e6))
            sys.stdout.flush()
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
