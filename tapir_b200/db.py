"""sqlite sink with the reference's schema and row semantics (tapir/db.py:15-61), plus a bulk writer.

Tables (byte-for-byte the reference's DDL): loci(id, locus), net(id, time, pi), discrete(id, time, pi),
interval(id, interval, pi, error).  Readers of this database (bin/tapir_plot.py, tapir_matplot.py,
tapir_compare.py, the README's SQL recipes) keep working unchanged."""
import os
import sqlite3

_DDL = (
    "CREATE TABLE loci (id INTEGER PRIMARY KEY AUTOINCREMENT, locus TEXT)",
    """CREATE TABLE net (id INT, time INT, pi FLOAT,
            FOREIGN KEY(id) REFERENCES loci(id) DEFERRABLE INITIALLY
            DEFERRED)""",
    """CREATE TABLE discrete (id INT, time INT, pi FLOAT, 
            FOREIGN KEY(id) REFERENCES loci(id) DEFERRABLE INITIALLY
            DEFERRED)""",
    """CREATE TABLE interval (id INT, interval TEXT, pi FLOAT,
            error FLOAT, FOREIGN KEY(id) REFERENCES loci(id) DEFERRABLE
            INITIALLY DEFERRED)""",
)


def create_probe_db(db_name, overwrite=None):
    """Open `db_name` and create the four tables -> (conn, cursor) (tapir/db.py:15-42).

    When the tables already exist the reference asks on stdin and, because of its always-true test,
    overwrites whatever the answer.  Here `overwrite=None` asks the same question and
    honours the answer; True/False skip the prompt."""
    conn = sqlite3.connect(db_name)
    c = conn.cursor()
    c.execute("PRAGMA foreign_keys = ON")
    try:
        for statement in _DDL:
            c.execute(statement)
    except sqlite3.OperationalError as err:
        if "table loci already exists" not in str(err):
            raise
        if overwrite is None:
            answer = input("\n\tPI database already exists.  Overwrite [Y/n]? ")
            overwrite = answer.strip().upper() in ("", "Y", "YES")
        if not overwrite:
            raise SystemExit(2)
        c.close()
        conn.close()
        os.remove(db_name)
        return create_probe_db(db_name, overwrite=True)
    return conn, c


def locus_name(path):
    """tapir/db.py:48: basename without its last extension."""
    return os.path.splitext(os.path.basename(path))[0]


def insert_pi_data(conn, c, pis):
    """One `loci` row, T `net` rows, one `discrete` row per --times entry and one `interval` row per
    epoch for every worker() 7-tuple (tapir/db.py:44-61), written with executemany."""
    for locus in pis:
        name, rates, mean_rate, pi, pi_net, times, epochs = locus
        c.execute("INSERT INTO loci(locus) VALUES (?)", (locus_name(name),))
        key = c.lastrowid
        c.executemany("INSERT INTO net VALUES (?,?,?)", ((key, k, float(v)) for k, v in enumerate(pi_net)))
        if times:
            c.executemany("INSERT INTO discrete VALUES (?,?,?)", ((key, int(k), float(v)) for k, v in times.items()))
        if epochs:
            c.executemany("INSERT INTO interval VALUES (?,?,?,?)",
                          ((key, k, v["sum(integral)"], v["sum(error)"]) for k, v in epochs.items()))
    return


_NET_SHARD_ROWS = 2000000          # below this many `net` rows the worker processes are not worth their start-up
_MAX_SHARDS = 8                    # SQLite attaches at most 10 databases; they cannot be detached before the commit


def _write_net_shard(job):
    """Worker process: one contiguous range of loci -> a database file holding their `net` rows (same DDL)."""
    path, keys, block = job
    T = block.shape[1]
    conn = sqlite3.connect(path)
    conn.execute("PRAGMA synchronous = OFF")
    conn.execute("PRAGMA journal_mode = OFF")
    conn.execute(_DDL[1])
    tt = list(range(T))
    step = max(1, 4000000 // max(T, 1))
    for l0 in range(0, len(keys), step):
        l1 = min(len(keys), l0 + step)
        ids = [k for k in keys[l0:l1] for _ in tt]
        conn.executemany("INSERT INTO net VALUES (?,?,?)", zip(ids, tt * (l1 - l0), block[l0:l1].reshape(-1).tolist()))
    conn.commit()
    conn.close()
    return path


def _insert_net_parallel(c, keys, pi_net, workers, tmp_dir):
    """The `net` table through worker processes: the Python side of executemany (a tuple and three bound values per
    row, ~1 us) is what limits the plain path, so contiguous ranges of loci are written to shard files in parallel and
    appended in order with INSERT ... SELECT at C speed.  Same rows, same order, same rowids.  The shards stay attached
    (as shard0 ...) until the caller has committed; detach_shards() then removes them."""
    import multiprocessing
    import numpy
    import tempfile
    n, T = pi_net.shape
    n_shards = min(_MAX_SHARDS, workers, max(1, (n * T) // (_NET_SHARD_ROWS // 2)))
    bounds = [n * k // n_shards for k in range(n_shards + 1)]
    work = tempfile.mkdtemp(prefix="tapir-net-", dir=tmp_dir)
    jobs = ((os.path.join(work, "shard%d.sqlite" % k), keys[bounds[k]:bounds[k + 1]],
             numpy.array(pi_net[bounds[k]:bounds[k + 1]]))           # a plain copy: the profile may sit in pinned memory
            for k in range(n_shards))
    with multiprocessing.get_context("fork").Pool(min(workers, n_shards)) as pool:
        for k, path in enumerate(pool.imap(_write_net_shard, jobs)):
            c.execute("ATTACH DATABASE ? AS shard%d" % k, (path,))
            c.execute("INSERT INTO net SELECT * FROM shard%d.net" % k)
            os.remove(path)                                           # stays open (attached) until it is detached
    try:
        os.rmdir(work)
    except OSError:
        pass
    return n_shards


def detach_shards(conn):
    """After the commit that follows insert_batch(..., workers > 1): detach the shard databases of the `net` table."""
    names = [row[1] for row in conn.execute("PRAGMA database_list").fetchall() if row[1].startswith("shard")]
    for name in names:
        conn.execute("DETACH DATABASE %s" % name)
    return len(names)


def insert_batch(conn, c, names, result, times, epochs, with_net=True, block=4096, workers=0, tmp_dir=None):
    """Bulk form of insert_pi_data straight from an engine.BatchResult (row l = locus l): same rows, same
    order, one executemany per table and block of loci, rows built from whole-array tolist() calls (no per-value
    Python arithmetic).  `with_net=False` skips the T-rows-per-locus `net` table.  `workers` > 1: a large `net`
    table is written through that many worker processes (_insert_net_parallel; call detach_shards(conn) after the
    commit).  With foreign-key enforcement off on the connection -- the ids are the ones generated here -- the shards
    are appended by SQLite's table-to-table transfer, four times faster again."""
    labels = ["{0}-{1}".format(s[0], s[1]) for s in epochs]
    n = len(names)
    c.executemany("INSERT INTO loci(locus) VALUES (?)", ((locus_name(name),) for name in names))
    # AUTOINCREMENT ids of one executemany are consecutive: the last one is the table's maximum
    last = c.execute("SELECT MAX(id) FROM loci").fetchone()[0] if n else 0
    keys = list(range(last - n + 1, last + 1))
    import numpy
    parallel = (with_net and result.pi_net is not None and n and workers > 1
                and n * result.pi_net.shape[1] >= _NET_SHARD_ROWS
                and "fork" in __import__("multiprocessing").get_all_start_methods())
    if parallel:
        _insert_net_parallel(c, keys, result.pi_net, workers, tmp_dir)
    elif with_net and result.pi_net is not None and n:
        T = result.pi_net.shape[1]
        tt = list(range(T))
        for l0 in range(0, n, max(1, block)):
            l1 = min(n, l0 + block)
            vals = numpy.asarray(result.pi_net[l0:l1]).reshape(-1).tolist()
            ids = numpy.repeat(numpy.asarray(keys[l0:l1]), T).tolist()
            c.executemany("INSERT INTO net VALUES (?,?,?)", zip(ids, tt * (l1 - l0), vals))
    if times and n:
        K = len(times)
        ids = numpy.repeat(numpy.asarray(keys), K).tolist()
        c.executemany("INSERT INTO discrete VALUES (?,?,?)",
                      zip(ids, [int(t) for t in times] * n, numpy.asarray(result.pi_times).reshape(-1).tolist()))
    if labels and n:
        E = len(labels)
        ids = numpy.repeat(numpy.asarray(keys), E).tolist()
        c.executemany("INSERT INTO interval VALUES (?,?,?,?)",
                      zip(ids, labels * n, numpy.asarray(result.integral).reshape(-1).tolist(),
                          numpy.asarray(result.abserr).reshape(-1).tolist()))
    return keys


def max_pi_epochs(conn):
    """README.rst:178-185: per locus the interval with the largest PI -> {interval: count of loci}."""
    rows = conn.execute("SELECT interval, COUNT(*) FROM (SELECT id, interval, MAX(pi) FROM interval GROUP BY id) "
                        "GROUP BY interval").fetchall()
    return dict(rows)
