"""Batched rate ingest: many HyPhy site-rate JSON files -> one CSR pack in pinned host memory, parsed by the
library's multi-threaded C++ reader (tapir_ingest_rates) instead of one json.load per file under the GIL.

It is the read half of parse_site_rates (tapir/compute.py:30-39); the files are NOT rewritten (the reference
appends a ``corrected_rates`` list to every file it reads, compute.py:40-43 -- use compute.parse_site_rates
when that side effect is wanted)."""
import ctypes
import os

import numpy as np

from . import _lib
from .engine import pinned_empty


def read_rate_files(paths, correction=1, threads=None, pinned=True):
    """-> (offsets int64[n+1], rates float64[nnz]) with rates = file rates / correction, files in the given order.
    Raises IOError for an unreadable file and ValueError for a malformed one."""
    lib = _lib.load()
    n = len(paths)
    arr = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    threads = int(threads or max(1, (os.cpu_count() or 2) - 1))
    offsets = np.zeros(n + 1, dtype=np.int64)
    needed = ctypes.c_int64(0)
    handle = ctypes.c_void_p()

    def check(rc):
        if rc != 0:
            msg = (lib.tapir_ingest_last_error() or b"?").decode("utf-8", "replace")
            raise (IOError if msg.startswith("Cannot open") else ValueError)(msg)

    # every file is read and parsed once: open keeps the rates, the buffer is sized from `needed`, copy fills it
    check(lib.tapir_ingest_open(arr, n, threads, ctypes.byref(handle), offsets.ctypes.data, ctypes.byref(needed)))
    try:
        alloc = pinned_empty if pinned else np.empty
        rates = alloc(int(needed.value), np.float64)
        if needed.value:
            check(lib.tapir_ingest_copy(handle, float(correction), rates.ctypes.data, int(needed.value)))
    finally:
        lib.tapir_ingest_close(handle)
    return offsets, rates
