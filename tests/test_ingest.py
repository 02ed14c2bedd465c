"""The library's multi-threaded rate-file reader (tapir_ingest_rates) against parse_site_rates' read half
(tapir/compute.py:30-39) on the reference's own fixture, on rewritten files, and on malformed input."""
import os
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "golden", "ref-data", "test-uniform-draw-weights.rates.json")


def test_matches_parse_site_rates(tmp_path, golden):
    from tapir_b200 import compute, ingest
    g, _ = golden
    paths = []
    for i in range(37):
        p = tmp_path / ("locus-%d.rates" % i)
        shutil.copyfile(SRC, p)
        paths.append(str(p))
    compute.parse_site_rates(paths[3], correction=100)                # a file the reference already rewrote
    before = [open(p).read() for p in paths]
    for threads in (1, 4):
        off, rates = ingest.read_rate_files(paths, correction=100, threads=threads, pinned=False)
        assert off.tolist() == [100 * i for i in range(38)]
        for i in range(37):
            assert np.array_equal(rates[off[i]:off[i + 1]], g["cfg1/rates"])      # bit-identical to numpy's rates/100
    assert [open(p).read() for p in paths] == before                  # nothing rewritten
    off, rates = ingest.read_rate_files([], pinned=False)
    assert off.tolist() == [0] and rates.size == 0


def test_json_corner_cases_and_errors(tmp_path):
    from tapir_b200 import ingest
    odd = tmp_path / "odd.rates"
    odd.write_text('{"x": [1, {"a": "q\\"uote}"}], "sites": {"freqs": {"A": 0.25}, "rates": [ {"rate": NaN, "site": 1},\n'
                   ' {"site":2,"rate":1e400}, {"rate": -Infinity}, {"rate": 3, "rate": 4e-2}, {"rate": 1e-400}, {"rate": 7} ],'
                   ' "corrected_rates": [{"site": 1, "rate": 9}] } }')
    off, rates = ingest.read_rate_files([str(odd)], correction=2, pinned=False)
    assert off.tolist() == [0, 6]
    assert np.isnan(rates[0]) and rates[1] == np.inf and rates[2] == -np.inf
    assert rates[3] == 0.02 and rates[4] == 0.0 and rates[5] == 3.5
    empty = tmp_path / "empty.rates"
    empty.write_text('{"sites": {"rates": []}}')
    off, rates = ingest.read_rate_files([str(empty), str(odd)], pinned=False)
    assert off.tolist() == [0, 0, 6]
    for text in ('{"sites": {"rates": [{"site": 1}]}}', '{"sites": {}}', '{"rates": []}', '{"sites": {"rates": [1,2]}}',
                 '{"sites": {"rates": [{"rate": 1}', ''):
        bad = tmp_path / "bad.rates"
        bad.write_text(text)
        with pytest.raises(ValueError):
            ingest.read_rate_files([str(empty), str(bad)], pinned=False)
    with pytest.raises(IOError):
        ingest.read_rate_files([str(tmp_path / "missing.rates")], pinned=False)


def test_truncated_and_trailing_documents_are_rejected(tmp_path):
    """json.load rejects a document that stops early or carries trailing data; so does the batched reader (it walks
    the document to its end instead of stopping when the rates list closes)."""
    import json
    import pytest
    from tapir_b200 import ingest
    good = {"sites": {"rates": [{"site": 1, "rate": 0.25, "ll": -1.0}, {"site": 2, "rate": 0.5}], "freqs": {"A": 0.25}},
            "note": "x\\"}
    text = json.dumps(good)
    ok = tmp_path / "ok.rates"
    ok.write_text(text)
    off, r = ingest.read_rate_files([str(ok)], pinned=False)
    assert list(off) == [0, 2] and list(r) == [0.25, 0.5]
    cut = text.index("]") + 1
    for k, bad_text in enumerate([text[:cut], text[:-1], text + " 1", text + "}", text[:cut] + "}", text[:-3]]):
        path = tmp_path / ("bad%d.rates" % k)
        path.write_text(bad_text)
        with pytest.raises(ValueError):
            json.loads(bad_text)
        with pytest.raises(ValueError):
            ingest.read_rate_files([str(ok), str(path)], pinned=False)
