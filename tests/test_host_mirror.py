"""Host side of the drop-in boundary (CPU only): argument helpers, tree scaling, rate-file ingest,
alignment culling mask and the sqlite sink, against the reference's own fixtures
(tests/golden/ref-data, copied verbatim from tapir/tests/test-data) and golden rows produced by the
reference's unmodified db.py (tests/golden/make_golden.py)."""
import argparse
import json
import os
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "golden", "ref-data")


# ---- tapir/tests/test_base.py:25-59 ----
def test_base_helpers(tmp_path):
    from tapir_b200 import base
    assert base.is_dir(str(tmp_path)) == str(tmp_path)
    with pytest.raises(argparse.ArgumentTypeError):
        base.is_dir(str(tmp_path / "missing"))
    assert base.get_output_type("plot.PDF") == "pdf"
    with pytest.raises(AssertionError):
        base.get_output_type("plot.txt")
    assert base.get_list_from_ints("1,2,3") == [1, 2, 3]
    with pytest.raises(argparse.ArgumentTypeError):
        base.get_list_from_ints("1,b")
    assert base.get_strings_from_items("a,b") == ["a", "b"]
    assert base.get_list_from_ranges("1-2,2-3,3-4") == [[1, 2], [2, 3], [3, 4]]
    with pytest.raises(argparse.ArgumentTypeError):
        base.get_list_from_ranges("1-x")
    for name in ("a.nex", "b.nexus", "c.rates"):
        (tmp_path / name).write_text("")
    assert sorted(map(os.path.basename, base.get_files(str(tmp_path), "*.nex,*.nexus"))) == ["a.nex", "b.nexus"]
    assert [os.path.basename(f) for f in base.get_files(str(tmp_path), "*.rates")] == ["c.rates"]
    with pytest.raises(IOError):
        base.get_files(str(tmp_path), "*.phy")
    m = tmp_path / "map.tsv"
    m.write_text("a.nex\t3\t40\n\nb.nex\t0\t9\n")
    assert dict(base.parse_subset_map_file(str(m))) == {"a.nex": [3, 40], "b.nex": [0, 9]}


def test_create_unique_dir(tmp_path):
    from tapir_b200 import base
    out = tmp_path / "out"
    out.mkdir()
    assert base.create_unique_dir(str(out)) == str(out)              # empty: used as is
    (out / "x").write_text("1")
    assert base.create_unique_dir(str(out)) == str(out) + ".1"       # 'out' exists -> out.1 created
    assert os.path.isdir(str(out) + ".1")
    assert base.create_unique_dir(str(out)) == str(out) + ".2"


# ---- tapir/tests/test_compute.py:76-91 ----
def test_tree_adjustment(tmp_path, golden):
    from tapir_b200 import compute, tree
    _, meta = golden
    depth, cf, path = compute.correct_branch_lengths(os.path.join(DATA, "Euteleost.tree"), "newick", d=str(tmp_path))
    assert (depth, cf) == (174.0, 100) and os.path.basename(path) == "Tree_100_174.0.newick"
    got, want = tree.read_tree(path), tree.parse_newick(meta["tree_expected_scaled"])
    assert [(n.label, n.length) for n in got.nodes()] == [(n.label, n.length) for n in want.nodes()]


def test_tree_semantics_and_formats(tmp_path):
    from tapir_b200 import compute, tree
    t = tree.parse_newick("((A:1,B:2.5)X:3,'C d':4[comment],(E,F:0):7)root;")
    assert t.depth() == 7.0 and t.length() == 17.5 and len(t.leaf_nodes()) == 5
    t.scale_edges(10)
    assert [n.length for n in t.leaf_nodes()] == [0.1, 0.25, 0.4, None, 0]      # falsy lengths untouched
    assert tree.parse_newick(t.to_newick()).to_newick() == t.to_newick()
    nex = tmp_path / "t.nex"
    nex.write_text("#NEXUS\nbegin trees;\n translate 1 danRer6, 2 'ory Lat';\n tree one = [&R] (1:174.0,2:100.0):0.0;\nend;\n")
    assert sorted(l.label for l in tree.read_tree(str(nex), "nexus").leaf_nodes()) == ["danRer6", "ory Lat"]
    # small mean branch length: correction factor stays 1 (compute.py:63-67)
    nw = tmp_path / "s.tree"
    nw.write_text("((a:1.2,b:0.7):0.4,c:2.0);")
    depth, cf, _ = compute.correct_branch_lengths(str(nw), "newick", d=str(tmp_path))
    assert cf == 1 and abs(depth - 2.0) < 1e-12
    with pytest.raises(ValueError):
        tree.parse_newick("((a,b);")


# ---- tapir/tests/test_compute.py:24-33 (made non-vacuous) ----
def test_parse_site_rates_and_rewrite(tmp_path, golden):
    from tapir_b200 import compute
    g, _ = golden
    src = os.path.join(DATA, "test-uniform-draw-weights.rates.json")
    assert np.array_equal(compute.parse_site_rates(src, test=True), g["fixture/test-parsed-rates"].reshape(-1))
    assert np.array_equal(compute.parse_site_rates(src, 10., test=True), g["bundled_rates_div10"])
    work = tmp_path / "locus.rates"
    shutil.copyfile(src, work)
    got = compute.parse_site_rates(str(work), correction=100)
    assert np.array_equal(got, g["cfg1/rates"])
    doc = json.load(open(work))                                       # rewritten with corrected_rates (compute.py:40-43)
    assert [r["site"] for r in doc["sites"]["corrected_rates"]] == list(range(1, 101))
    assert np.array_equal([r["rate"] for r in doc["sites"]["corrected_rates"]], g["cfg1/rates"])
    assert np.array_equal(compute.parse_site_rates(str(work), test=True), g["bundled_rates"])   # raw rates untouched
    with pytest.raises(IOError):
        compute.parse_site_rates(str(tmp_path / "absent.rates"), test=True)


# ---- tapir/tests/test_compute.py:93-121 (made non-vacuous) ----
def test_informative_sites_and_culling(golden):
    from tapir_b200 import compute
    g, _ = golden
    small = compute.get_informative_sites(os.path.join(DATA, "informativeness_cutoff.nex"), 3)
    assert np.array_equal(small, [np.nan, np.nan, 1.0, 1.0], equal_nan=True)
    full = compute.get_informative_sites(os.path.join(DATA, "chr1_918.nex"), 3)
    assert np.array_equal(full, g["fixture/chr1_918-test-cutoff-values"], equal_nan=True)
    rates = compute.parse_site_rates(os.path.join(DATA, "test-uniform-draw-weights.rates.json"), 10., test=True)
    culled = compute.cull_uninformative_rates(rates, full[:100])
    assert np.array_equal(culled, g["fixture/test-culled-rates"], equal_nan=True)


def test_get_time():
    from tapir_b200 import compute
    t = compute.get_time(0, 174)
    assert t.shape == (174, 1) and t.dtype.kind == "i" and t[0, 0] == 0 and t[-1, 0] == 173
    assert compute.get_time(2, 11, 3).reshape(-1).tolist() == [2, 5, 8]


# ---- tapir/db.py ----
def _db_golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_db_golden.json")) as fh:
        return json.load(fh)


def _cfg1_tuple(golden):
    g, meta = golden
    times = dict(zip(meta["times"], g["cfg1/pi_times"].tolist()))
    epochs = {"{0}-{1}".format(a, b): {"sum(integral)": float(i), "sum(error)": float(e)}
              for (a, b), i, e in zip(meta["epochs"], g["cfg1/integral"], g["cfg1/error"])}
    return ("/some/dir/locus-one.rates", g["cfg1/rates"], float(g["cfg1/mean_rate"]), None, g["cfg1/pi_net"], times, epochs)


def _dump(conn):
    return {"ddl": [r[0] for r in conn.execute("SELECT sql FROM sqlite_master WHERE type='table' ORDER BY name")],
            "loci": [list(r) for r in conn.execute("SELECT * FROM loci")],
            "net": [list(r) for r in conn.execute("SELECT * FROM net ORDER BY time")],
            "discrete": [list(r) for r in conn.execute("SELECT * FROM discrete ORDER BY time")],
            "interval": [list(r) for r in conn.execute("SELECT * FROM interval ORDER BY interval")]}


def test_db_schema_and_rows_match_reference(tmp_path, golden):
    from tapir_b200 import db
    want = _db_golden()
    conn, c = db.create_probe_db(str(tmp_path / "pi.sqlite"))
    db.insert_pi_data(conn, c, [_cfg1_tuple(golden)])
    conn.commit()
    assert _dump(conn) == want                                       # DDL text and every row, bit for bit
    assert db.max_pi_epochs(conn) == {"164-174": 1}                  # README.rst:178-185
    conn.close()
    with pytest.raises(SystemExit):
        db.create_probe_db(str(tmp_path / "pi.sqlite"), overwrite=False)
    conn, c = db.create_probe_db(str(tmp_path / "pi.sqlite"), overwrite=True)
    assert conn.execute("SELECT COUNT(*) FROM loci").fetchone()[0] == 0
    conn.close()


def test_db_bulk_writer_equals_row_writer(tmp_path, golden):
    from tapir_b200 import db
    from tapir_b200.engine import BatchResult
    g, meta = golden
    res = BatchResult(pi_net=g["cfg1/pi_net"][None, :], pi_times=g["cfg1/pi_times"][None, :],
                      integral=g["cfg1/integral"][None, :], abserr=g["cfg1/error"][None, :], max_pi=None, argmax_t=None,
                      argmax_epoch=None, mean_rate=None, n_finite=None, n_multi_panel=None)
    conn, c = db.create_probe_db(str(tmp_path / "bulk.sqlite"))
    db.insert_batch(conn, c, ["/some/dir/locus-one.rates"], res, meta["times"], meta["epochs"])
    conn.commit()
    assert _dump(conn) == _db_golden()
    conn.close()


def test_cli_arguments_match_reference(tmp_path):
    from tapir_b200 import driver
    (tmp_path / "in").mkdir()
    (tmp_path / "out").mkdir()
    args = driver.get_args([str(tmp_path / "in"), "some.tree", "--output", str(tmp_path / "out"),
                            "--intervals=32-42,88-98,95-105,164-174", "--times=37,93,100,170", "--site-rates",
                            "--multiprocessing", "--tree-format", "nexus", "--threshold", "4"])
    assert args.times == [37, 93, 100, 170] and args.intervals == [[32, 42], [88, 98], [95, 105], [164, 174]]
    assert args.site_rates and args.multiprocessing and args.tree_format == "nexus" and args.threshold == 4
    assert os.path.isabs(args.tree) and args.hyphy == "hyphy2" and args.template is None
    with pytest.raises(SystemExit):
        driver.get_args([str(tmp_path / "in"), "some.tree", "--times=1"])       # --intervals is required


def test_load_rates_hyphy_mode_with_stub_binary(tmp_path):
    """HyPhy stays an external step: a stub 'hyphy' that writes the rate JSON exercises the Popen protocol
    (alignment, tree, output path on stdin; bin/tapir_compute.py:90-102), the culling mask and the subset slice."""
    from tapir_b200 import driver
    src = os.path.join(DATA, "test-uniform-draw-weights.rates.json")
    stub = tmp_path / "hyphy_stub.py"
    stub.write_text("#!/usr/bin/env python\nimport sys, shutil\nlines = sys.stdin.read().split('\\n')\n"
                    "shutil.copyfile(%r, lines[2])\nprint('done')\n" % src)
    stub.chmod(0o755)
    from tapir_b200 import compute
    rows = compute._read_nexus_matrix(os.path.join(DATA, "chr1_918.nex"))       # 5 taxa x 226; the rate file has 100 sites
    aln = str(tmp_path / "chr1_918.nex")
    with open(aln, "w") as fh:
        fh.write("#NEXUS\nbegin data;\n dimensions ntax=5 nchar=100;\n format datatype=dna missing=? gap=-;\nmatrix\n")
        fh.write("".join("t%d %s\n" % (i, r[:100]) for i, r in enumerate(rows)) + ";\nend;\n")
    out = str(tmp_path / "chr1_918.nex.rates")
    towrite = "\n".join([aln, "tree.newick", out])
    params = [None, str(stub), "template.bf", towrite, out, 10.0, aln, [1], [[1, 2]], 3, {"chr1_918.nex": [5, 60]}]
    name, rates = driver.load_rates(params)
    assert name == aln and rates.shape == (55,)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    assert np.array_equal(rates, golden["fixture/test-culled-rates"][5:60], equal_nan=True)
    with pytest.raises(Exception, match="hyphy not found"):
        driver.load_rates([None, str(tmp_path / "nope"), "t", towrite, out, 1, aln, [], [], 3, {}])


def test_bench_reads_profile_kernel_traffic_from_the_committed_ncu_summary():
    """bench.py's roofline.traffic comes from profiles/*_ncu_full_cfg2.txt; the block of the shipped profile kernel
    (profile_mom_kernel, or profile_kernel in older captures) must be found."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    t = bench.profile_traffic("cfg2")
    assert t is not None and 2.0e7 < t < 3.0e7        # 20 MB of rates read once, under 1 MB written (L2-resident profile)
    u = bench.profile_ncu_utilisation("cfg5")         # the same capture's unit utilisations ride in the roofline object
    assert u["source"].startswith("profiles/") and 0 < u["dmma_subpipe_pct"] < 100 and 0 < u["l1_data_pipe_pct"] < 100
    assert u["instruction_cache_hit_pct"] > 90


def test_rate_file_rewrite_is_json_dump_indent4_byte_for_byte(tmp_path):
    """parse_site_rates rewrites every rate file as the reference does (json.dump(data, handle, indent=4),
    compute.py:41-43); the writer used here is a faster one that must produce the same bytes for anything a rate file
    can hold, and hand over to json.dumps for anything else."""
    import json
    from tapir_b200.compute import dumps_indent4, parse_site_rates
    rng = np.random.default_rng(9)
    lam = rng.gamma(0.5, 4.0, 300)
    doc = {"sites": {"rates": [{"site": k + 1, "subst": 0.0, "rate": round(float(v), 4), "ll": -1.5e-7} for k, v in enumerate(lam)]},
           "meta": {"file": "aé\n\"b\"", "n": 300, "ok": True, "none": None, "empty": [], "nested": [[1, 2.5], {}, []]}}
    cases = [doc,
             {"a": [{"x": 1, "y": float("nan")}] * 9, "b": [{"x": 1.0, "y": float("inf")}, {"x": 2.5, "y": -float("inf")}] * 5},
             [{"k%": i, "s": "q\\", "n": None, "t": True} for i in range(10)],
             [{"a": 1}] * 8 + [{"b": 2}], [{"a": 1, "b": 2}] * 8 + [{"b": 2, "a": 1}], [{"a": 1}] * 7 + [{"a": 1.5}],
             {"f": -0.0, "g": 1e-300, "h": 1e22, "i": 12345678901234567890}, [{"a": [1]}] * 9, [{}] * 9, [{1: 2}] * 9,
             (1, 2, 3), "str", 3.5, None, [{"a": True, "b": None}] * 9, [{"a": 1, "b": True}] * 8 + [{"a": 1, "b": None}]]
    for c in cases:
        assert dumps_indent4(c) == json.dumps(c, indent=4)
    path = tmp_path / "locus.rates"
    path.write_text(json.dumps(doc))
    got = parse_site_rates(str(path), correction=100)
    want = json.loads(json.dumps(doc))
    want["sites"]["corrected_rates"] = [{"site": k + 1, "rate": v} for k, v in enumerate((np.array([r["rate"] for r in doc["sites"]["rates"]]) / 100).tolist())]
    assert path.read_text() == json.dumps(want, indent=4)
    assert np.array_equal(got, np.array([r["rate"] for r in doc["sites"]["rates"]]) / 100)


def test_net_table_through_worker_processes_is_row_identical(tmp_path, monkeypatch):
    """db.insert_batch(..., workers > 1) writes a large `net` table through shard databases built by worker processes
    and appended in order: same rows, same rowids, same other tables as the plain bulk writer -- with foreign keys on
    (row-by-row append) and off (SQLite's table-to-table transfer); the shards are detached after the commit."""
    from tapir_b200 import db

    class Result:
        pass
    rng = np.random.default_rng(4)
    n, T = 300, 200
    r = Result()
    r.pi_net, r.pi_times = rng.random((n, T)), rng.random((n, 3))
    r.integral, r.abserr = rng.random((n, 2)), rng.random((n, 2))
    names = ["dir/locus-%04d.rates" % i for i in range(n)]
    monkeypatch.setattr(db, "_NET_SHARD_ROWS", 10000)
    tables = {}
    for label, workers, fk in (("plain", 0, True), ("workers", 5, True), ("workers, no fk checks", 5, False)):
        d = tmp_path / label.replace(" ", "_").replace(",", "")
        d.mkdir()
        conn, c = db.create_probe_db(str(d / "pi.sqlite"), overwrite=True)
        if not fk:
            c.execute("PRAGMA foreign_keys = OFF")
        keys = db.insert_batch(conn, c, names, r, [5, 17, 199], [[32, 42], [88, 98]], workers=workers, tmp_dir=str(d))
        conn.commit()
        assert db.detach_shards(conn) == (0 if workers == 0 else 5)
        assert [f for f in os.listdir(d) if f != "pi.sqlite"] == []
        assert keys == list(range(1, n + 1))
        tables[label] = {t: conn.execute("SELECT rowid, * FROM %s ORDER BY rowid" % t).fetchall()
                         for t in ("loci", "net", "discrete", "interval")}
        if not fk:
            c.execute("PRAGMA foreign_keys = ON")
            assert conn.execute("PRAGMA foreign_key_check").fetchall() == []
        conn.close()
    assert len(tables["plain"]["net"]) == n * T
    assert tables["plain"] == tables["workers"] == tables["workers, no fk checks"]
