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


def insert_batch(conn, c, names, result, times, epochs, with_net=True, block=4096):
    """Bulk form of insert_pi_data straight from an engine.BatchResult (row l = locus l): same rows, same
    order, one executemany per table and block of loci, rows built from whole-array tolist() calls (no per-value
    Python arithmetic).  `with_net=False` skips the T-rows-per-locus `net` table."""
    labels = ["{0}-{1}".format(s[0], s[1]) for s in epochs]
    n = len(names)
    c.executemany("INSERT INTO loci(locus) VALUES (?)", ((locus_name(name),) for name in names))
    # AUTOINCREMENT ids of one executemany are consecutive: the last one is the table's maximum
    last = c.execute("SELECT MAX(id) FROM loci").fetchone()[0] if n else 0
    keys = list(range(last - n + 1, last + 1))
    import numpy
    if with_net and result.pi_net is not None and n:
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
