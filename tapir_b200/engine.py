"""Batched multi-locus driver over libtapir_cuda (the GPU replacement for mapping
``worker()`` over loci, bin/tapir_compute.py:159-164).

Many loci are packed into one CSR pair ``(offsets int64[L+1], rates float64[nnz])`` in pinned
host memory and handed to the C ABI in one call.  numpy is used for host arrays only; all
arithmetic of the hot path happens in the CUDA library.  There is no CPU fallback.
"""
import ctypes
import weakref
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Problem, Results, Timing, check


@dataclass
class BatchResult:
    """Per-locus outputs of one batch (row l = locus l of the CSR pack)."""
    pi_net: Optional[np.ndarray]   # (L, T) net PI profile, None unless requested
    pi_times: np.ndarray           # (L, K) profile at --times
    integral: np.ndarray           # (L, E) sum(integral) per epoch
    abserr: np.ndarray             # (L, E) sum(error) per epoch
    max_pi: np.ndarray             # (L,)
    argmax_t: np.ndarray           # (L,) int64
    argmax_epoch: np.ndarray       # (L,) int32, index into the epoch list (README.rst:178-185)
    mean_rate: np.ndarray          # (L,)
    n_finite: np.ndarray           # (L,) int64
    n_multi_panel: np.ndarray      # (L,) int64


def _free_pinned(lib, ptr):
    lib.tapir_cuda_host_free(ptr)


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory from tapir_cuda_host_alloc."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    nbytes = max(1, n * dtype.itemsize)
    ptr = ctypes.c_void_p()
    check(lib.tapir_cuda_host_alloc(ctypes.byref(ptr), nbytes))
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    weakref.finalize(buf, _free_pinned, lib, ptr.value)
    return arr


def pack_csr(loci: Sequence[np.ndarray], pinned: bool = True):
    """Pack per-locus rate vectors into (offsets, rates).  NaN entries mark culled sites."""
    lens = np.fromiter((len(x) for x in loci), dtype=np.int64, count=len(loci))
    alloc = pinned_empty if pinned else np.empty
    offsets = alloc(len(loci) + 1, np.int64)
    offsets[0] = 0
    np.cumsum(lens, out=offsets[1:])
    rates = alloc(int(offsets[-1]), np.float64)
    if len(loci):
        np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in loci], out=rates)
    return offsets, rates


def shard_csr(offsets, rates, index, pinned: bool = False):
    """CSR slice holding the loci `index` (ascending) -> (offsets_p, rates_p), without a Python loop over loci: a run
    of consecutive loci is one slice copy, anything else one gather through a site index."""
    offsets = np.asarray(offsets, dtype=np.int64)
    index = np.asarray(index, dtype=np.int64)
    alloc = pinned_empty if pinned else np.empty
    lens = offsets[index + 1] - offsets[index]
    off_p = alloc(index.size + 1, np.int64)
    off_p[0] = 0
    np.cumsum(lens, out=off_p[1:])
    rates_p = alloc(int(off_p[-1]), np.float64)
    if index.size and int(index[-1]) - int(index[0]) + 1 == index.size:        # contiguous block of loci
        rates_p[:] = rates[offsets[index[0]]:offsets[index[-1] + 1]]
    elif index.size:
        src = np.repeat(offsets[index] - off_p[:-1], lens) + np.arange(int(off_p[-1]), dtype=np.int64)
        np.take(rates, src, out=rates_p)
    return off_p, rates_p


def partition_loci(site_counts: np.ndarray, n_parts: int):
    """Size-balanced locus partition (longest-processing-time greedy, ties by index): returns a
    list of index arrays, one per part, each sorted ascending so a part's CSR slice keeps the
    original locus order."""
    import heapq
    site_counts = np.asarray(site_counts, dtype=np.int64)
    n = len(site_counts)
    if n and (site_counts == site_counts[0]).all():
        # uniform loci: contiguous blocks, no need for the greedy pass
        bounds = np.linspace(0, n, n_parts + 1).astype(np.int64)
        return [np.arange(bounds[i], bounds[i + 1], dtype=np.int64) for i in range(n_parts)]
    order = np.argsort(-site_counts, kind="stable")
    owner = np.empty(n, dtype=np.int64)
    heap = [(0, p) for p in range(n_parts)]          # (load, part): the least loaded part, lowest index on ties
    costs = (site_counts[order] + 1).tolist()
    for k, cost in enumerate(costs):
        load, p = heap[0]
        owner[k] = p
        heapq.heapreplace(heap, (load + cost, p))
    part_of = np.empty(n, dtype=np.int64)
    part_of[order] = owner
    return [np.nonzero(part_of == p)[0].astype(np.int64) for p in range(n_parts)]


class Engine:
    """One CUDA context on one device."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        ctx = ctypes.c_void_p()
        check(self._lib.tapir_cuda_create(device, ctypes.byref(ctx)))
        self._ctx = ctx
        self.device = device
        self._resident = None
        self._finalizer = weakref.finalize(self, self._lib.tapir_cuda_destroy, ctx)

    def close(self):
        self._finalizer()

    # ---- helpers ----
    @staticmethod
    def _prep(offsets, rates, n_grid, times, epochs, want_profile, flags):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        rates = np.ascontiguousarray(rates, dtype=np.float64)
        time_idx = np.ascontiguousarray(np.asarray(times, dtype=np.int64).reshape(-1))
        epoch_ab = np.ascontiguousarray(np.asarray(epochs, dtype=np.float64).reshape(-1))
        if epoch_ab.size % 2:
            raise ValueError("epochs must be (start, stop) pairs")
        p = Problem()
        p.n_loci = offsets.size - 1
        p.offsets = offsets.ctypes.data
        p.rates = rates.ctypes.data if rates.size else None
        p.n_grid = int(n_grid)
        p.n_times = time_idx.size
        p.time_idx = time_idx.ctypes.data if time_idx.size else None
        p.n_epochs = epoch_ab.size // 2
        p.epoch_ab = epoch_ab.ctypes.data if epoch_ab.size else None
        p.flags = int(flags) | (_lib.WANT_PROFILE if want_profile else 0)
        return p, (offsets, rates, time_idx, epoch_ab)

    @staticmethod
    def _alloc_results(L, T, K, E, want_profile, pinned):
        alloc = pinned_empty if pinned else np.empty
        res = BatchResult(
            pi_net=alloc((L, T), np.float64) if want_profile else None,
            pi_times=alloc((L, K), np.float64),
            integral=alloc((L, E), np.float64),
            abserr=alloc((L, E), np.float64),
            max_pi=alloc((L,), np.float64),
            argmax_t=alloc((L,), np.int64),
            argmax_epoch=alloc((L,), np.int32),
            mean_rate=alloc((L,), np.float64),
            n_finite=alloc((L,), np.int64),
            n_multi_panel=alloc((L,), np.int64),
        )
        r = Results()
        for name in _lib.RESULT_FIELDS:
            arr = getattr(res, name)
            setattr(r, name, arr.ctypes.data if arr is not None and arr.size else None)
        return res, r

    # ---- whole path, host buffers in / host buffers out ----
    def run(self, offsets, rates, n_grid, times, epochs, want_profile=True, flags=0, out: BatchResult = None,
            pinned_results=False, fp32=False):
        """H2D + kernels + D2H for one CSR batch.  Raises TapirCudaError on any failure.
        fp32: TAPIR_FP32_MODE -- the profile (and the slow sites' epoch sums) in single precision, ~1e-6 relative."""
        p, keep = self._prep(offsets, rates, n_grid, times, epochs, want_profile, flags | (_lib.FP32_MODE if fp32 else 0))
        if out is None:
            res, r = self._alloc_results(p.n_loci, p.n_grid, p.n_times, p.n_epochs, want_profile, pinned_results)
        else:
            res, r = out, Results()
            for name in _lib.RESULT_FIELDS:
                arr = getattr(res, name)
                setattr(r, name, arr.ctypes.data if arr is not None and arr.size else None)
        check(self._lib.tapir_cuda_run(self._ctx, ctypes.byref(p), ctypes.byref(r)), self._ctx)
        del keep
        return res

    # ---- staged path ----
    def upload(self, offsets, rates, n_grid, times, epochs, want_profile=True, flags=0, fp32=False):
        p, keep = self._prep(offsets, rates, n_grid, times, epochs, want_profile, flags | (_lib.FP32_MODE if fp32 else 0))
        check(self._lib.tapir_cuda_upload(self._ctx, ctypes.byref(p)), self._ctx)
        self._resident = (p.n_loci, p.n_grid, p.n_times, p.n_epochs, want_profile)
        del keep

    def compute(self) -> Timing:
        t = Timing()
        check(self._lib.tapir_cuda_compute(self._ctx, ctypes.byref(t)), self._ctx)
        return t

    def download(self, out: BatchResult = None, pinned_results=False) -> BatchResult:
        if self._resident is None:
            raise _lib.TapirCudaError(_lib.TAPIR_E_STATE, "download before upload")
        L, T, K, E, want_profile = self._resident
        if out is None:
            res, r = self._alloc_results(L, T, K, E, want_profile, pinned_results)
        else:
            res, r = out, Results()
            for name in _lib.RESULT_FIELDS:
                arr = getattr(res, name)
                setattr(r, name, arr.ctypes.data if arr is not None and arr.size else None)
        check(self._lib.tapir_cuda_download(self._ctx, ctypes.byref(r)), self._ctx)
        return res

    def device_results(self) -> Results:
        r = Results()
        check(self._lib.tapir_cuda_device_results(self._ctx, ctypes.byref(r)), self._ctx)
        return r

    SMALL_SECTIONS = ("pi_times", "integral", "abserr", "max_pi", "argmax_t", "mean_rate", "n_finite", "n_multi_panel",
                      "argmax_epoch")

    def device_small_block(self):
        """(device pointer, bytes, {section: byte offset}) of the block holding every small per-locus vector."""
        ptr, nbytes, offs = ctypes.c_void_p(), ctypes.c_size_t(), (ctypes.c_size_t * 9)()
        check(self._lib.tapir_cuda_device_small_block(self._ctx, ctypes.byref(ptr), ctypes.byref(nbytes),
                                                      ctypes.byref(offs)), self._ctx)
        return ptr.value, nbytes.value, dict(zip(self.SMALL_SECTIONS, list(offs)))

    # ---- ranking queries on the resident results ----
    def rank(self, n_top: int):
        """(top_locus[E,n_top], top_pi[E,n_top], epoch_count[E]) of the batch last uploaded and computed."""
        if self._resident is None:
            raise _lib.TapirCudaError(_lib.TAPIR_E_STATE, "rank before upload")
        E = self._resident[3]
        top_locus = np.full((E, n_top), -1, dtype=np.int64)
        top_pi = np.full((E, n_top), np.nan)
        count = np.zeros(E, dtype=np.int64)
        check(self._lib.tapir_cuda_rank(self._ctx, n_top, top_locus.ctypes.data if top_locus.size else None,
                                        top_pi.ctypes.data if top_pi.size else None,
                                        count.ctypes.data if E else None), self._ctx)
        return top_locus, top_pi, count

    # ---- NVLink push gather (CUDA IPC peer memory on the owner's GPU) ----
    def gather_create(self, slot_bytes: int, n_slots: int) -> bytes:
        """Owner (rank 0): allocate the gather buffer, return the 64-byte IPC handle for the other ranks."""
        handle = ctypes.create_string_buffer(64)
        check(self._lib.tapir_cuda_gather_create(self._ctx, slot_bytes, n_slots, handle), self._ctx)
        return handle.raw

    def gather_attach(self, handle, slot_bytes: int, slot: int):
        """Every rank (owner: handle=None): from now on compute() ends by pushing the small block into `slot`."""
        buf = ctypes.create_string_buffer(handle, 64) if handle is not None else None
        check(self._lib.tapir_cuda_gather_attach(self._ctx, buf, slot_bytes, slot), self._ctx)

    def gather_read(self, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        check(self._lib.tapir_cuda_gather_read(self._ctx, out.ctypes.data, nbytes), self._ctx)
        return out

    def gather_detach(self):
        check(self._lib.tapir_cuda_gather_detach(self._ctx), self._ctx)

    # ---- element-wise parity API ----
    def pi_matrix(self, time, rates):
        time = np.ascontiguousarray(np.asarray(time, dtype=np.float64).reshape(-1))
        rates = np.ascontiguousarray(np.asarray(rates, dtype=np.float64).reshape(-1))
        out = np.empty((time.size, rates.size), dtype=np.float64)
        check(self._lib.tapir_cuda_pi_matrix(self._ctx, time.ctypes.data, time.size, rates.ctypes.data, rates.size,
                                             out.ctypes.data), self._ctx)
        return out

    def nansum_rows(self, matrix):
        matrix = np.ascontiguousarray(matrix, dtype=np.float64)
        if matrix.ndim != 2:
            raise ValueError("nansum_rows expects a 2-D array")
        out = np.empty(matrix.shape[0], dtype=np.float64)
        check(self._lib.tapir_cuda_nansum_rows(self._ctx, matrix.ctypes.data if matrix.size else None, matrix.shape[0],
                                               matrix.shape[1], out.ctypes.data), self._ctx)
        return out

    def site_integrals(self, a, b, rates):
        rates = np.ascontiguousarray(np.asarray(rates, dtype=np.float64).reshape(-1))
        integ = np.empty(rates.size)
        err = np.empty(rates.size)
        panels = np.empty(rates.size, dtype=np.int32)
        check(self._lib.tapir_cuda_site_integrals(self._ctx, float(a), float(b), rates.ctypes.data, rates.size,
                                                  integ.ctypes.data, err.ctypes.data, panels.ctypes.data), self._ctx)
        return integ, err, panels

    def run_fn(self, offsets, rates, n_grid, times, epochs):
        """compute_fn for sharded.run_sharded: one CSR slice -> BatchResult with the profile."""
        return self.run(offsets, rates, n_grid, times, epochs, want_profile=True)

    # ---- measurement helpers ----
    def set_serial_kernels(self, serial: bool):
        check(self._lib.tapir_cuda_set_serial_kernels(self._ctx, 1 if serial else 0), self._ctx)

    def transfer_bytes(self):
        h2d, d2h = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._lib.tapir_cuda_transfer_bytes(self._ctx, ctypes.byref(h2d), ctypes.byref(d2h)), self._ctx)
        return h2d.value, d2h.value

    def fp64_peak(self):
        tf, ms = ctypes.c_double(), ctypes.c_float()
        check(self._lib.tapir_cuda_fp64_peak(self._ctx, ctypes.byref(tf), ctypes.byref(ms)), self._ctx)
        return tf.value, ms.value

    def flush_l2(self):
        check(self._lib.tapir_cuda_flush_l2(self._ctx), self._ctx)


def bind_host_thread_near(device: int) -> bool:
    """Pin the calling thread to the CPUs NVML reports as closest to `device` (so that pinned buffers allocated
    afterwards are first-touched on the GPU's NUMA node).  Best effort: False when NVML is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        return True
    except Exception:
        return False


def device_count() -> int:
    n = ctypes.c_int()
    lib = _lib.load()
    rc = lib.tapir_cuda_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0
