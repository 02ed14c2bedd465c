"""CPU-only checks of the boundary: the library loads and exports exactly what the header declares
(no compute call is made here -- there is no GPU in the authoring container)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tapir_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tapir_(?:cuda|ingest|mom)_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from tapir_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SYMBOLS) == names          # the ctypes table covers the whole header
    assert lib.tapir_cuda_version() == 2


def test_struct_layouts_match_header():
    from tapir_b200 import _lib
    assert ctypes.sizeof(_lib.Problem) == 72
    assert ctypes.sizeof(_lib.Results) == 80
    assert ctypes.sizeof(_lib.Timing) == 40
    assert ctypes.sizeof(_lib.MomEpoch) == 312


def test_no_device_fails_loudly():
    """Without a GPU the product path must raise, never fall back to a CPU implementation."""
    from tapir_b200 import _lib, engine
    if engine.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.TapirCudaError):
        engine.Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tapir_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, os.path.join(dirpath, f)


def test_partition_loci_balanced():
    from tapir_b200.engine import partition_loci
    rng = np.random.default_rng(1)
    s = rng.integers(200, 5000, size=1000)
    parts = partition_loci(s, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(1000))
    loads = np.array([s[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01
    parts = partition_loci(np.full(10, 800), 4)
    assert [len(p) for p in parts] == [2, 3, 2, 3] or sum(len(p) for p in parts) == 10
