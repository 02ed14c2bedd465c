"""Locus-sharded execution over the GPUs of one box: one process per GPU (torch.distributed), every rank
computes a size-balanced subset of the loci independently, and only the small per-locus result vectors
are gathered to rank 0 (the reference's equivalent is Pool.map over loci, bin/tapir_compute.py:159-164).

torch is used for the process group only (NCCL over NVLink on GPUs, gloo in the CPU tests); the
arithmetic is whatever ``compute_fn`` does -- Engine.run in the product; the gloo tests pass a CPU stand-in.
"""
import numpy as np

from .engine import BatchResult, partition_loci, shard_csr  # noqa: F401 (shard_csr re-exported)

SMALL_FIELDS = ("pi_times", "integral", "abserr", "max_pi", "argmax_t", "argmax_epoch", "mean_rate", "n_finite",
                "n_multi_panel")


def pack_small(res: BatchResult):
    """All per-locus vectors of a BatchResult as one float64 matrix [L, K + 2E + 6] (integers < 2^53 exact)."""
    cols = [res.pi_times, res.integral, res.abserr, res.max_pi[:, None], res.argmax_t[:, None].astype(np.float64),
            res.argmax_epoch[:, None].astype(np.float64), res.mean_rate[:, None], res.n_finite[:, None].astype(np.float64),
            res.n_multi_panel[:, None].astype(np.float64)]
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def unpack_small(m, K, E, pi_net=None):
    o = K + 2 * E
    return BatchResult(pi_net=pi_net, pi_times=m[:, :K].copy(), integral=m[:, K:K + E].copy(),
                       abserr=m[:, K + E:o].copy(), max_pi=m[:, o].copy(), argmax_t=m[:, o + 1].astype(np.int64),
                       argmax_epoch=m[:, o + 2].astype(np.int32), mean_rate=m[:, o + 3].copy(),
                       n_finite=m[:, o + 4].astype(np.int64), n_multi_panel=m[:, o + 5].astype(np.int64))


def run_sharded(compute_fn, offsets, rates, n_grid, times, epochs, group=None, gather_profile=False, device=None):
    """Every rank calls this with the SAME full CSR batch.  Rank r computes partition r with
    ``compute_fn(offsets_p, rates_p, n_grid, times, epochs) -> BatchResult``; rank 0 returns the BatchResult
    of the whole batch in the original locus order (db ids follow that order, tapir/db.py:47-49), the
    other ranks return None.  Without torch.distributed (group is None and no default group) it is a
    plain single-device call."""
    import torch
    import torch.distributed as dist

    offsets = np.asarray(offsets, dtype=np.int64)
    L = offsets.size - 1
    K, E = len(times), len(epochs)
    if not (dist.is_available() and dist.is_initialized()):
        return compute_fn(offsets, rates, n_grid, times, epochs)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition_loci(offsets[1:] - offsets[:-1], world)
    mine = parts[rank]
    off_p, rates_p = shard_csr(offsets, rates, mine)
    res = compute_fn(off_p, rates_p, n_grid, times, epochs)

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" and device is None else (device or "cpu")
    width = K + 2 * E + 6
    pad = max(len(p) for p in parts)

    def gather(local, ncols):
        buf = torch.zeros((pad, ncols), dtype=torch.float64, device=dev)
        if local.shape[0]:
            buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
        dst = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, dst, dst=0, group=group)
        return [d.cpu().numpy() for d in dst] if rank == 0 else None

    small = gather(pack_small(res), width)
    prof = gather(res.pi_net, int(n_grid)) if gather_profile and res.pi_net is not None else None
    if rank != 0:
        return None
    merged = np.empty((L, width), dtype=np.float64)
    pi_net = np.empty((L, int(n_grid)), dtype=np.float64) if prof is not None else None
    for r, idx in enumerate(parts):            # un-permute to the original locus order
        merged[idx] = small[r][:idx.size]
        if pi_net is not None:
            pi_net[idx] = prof[r][:idx.size]
    return unpack_small(merged, K, E, pi_net)


def run_threaded(offsets, rates, n_grid, times, epochs, devices, want_profile=True):
    """Single-process form for the command line: one host thread and one Engine per device (ctypes releases
    the GIL during the call), loci partitioned as in run_sharded, results merged in the original order.
    No collective: every device copies its own results to the host."""
    import threading

    from .engine import Engine

    offsets = np.asarray(offsets, dtype=np.int64)
    L = offsets.size - 1
    devices = list(devices)
    if len(devices) <= 1 or L < 2 * len(devices):
        eng = Engine(devices[0] if devices else 0)
        try:
            return eng.run(offsets, rates, n_grid, times, epochs, want_profile=want_profile)
        finally:
            eng.close()
    parts = partition_loci(offsets[1:] - offsets[:-1], len(devices))
    results, errors = [None] * len(devices), [None] * len(devices)

    def work(k):
        try:
            eng = Engine(devices[k])
            try:
                off_p, rates_p = shard_csr(offsets, rates, parts[k])
                results[k] = eng.run(off_p, rates_p, n_grid, times, epochs, want_profile=want_profile)
            finally:
                eng.close()
        except Exception as err:          # noqa: BLE001 - re-raised in the caller's thread
            errors[k] = err

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for err in errors:
        if err is not None:
            raise err
    K, E = len(times), len(epochs)
    merged = np.empty((L, K + 2 * E + 6), dtype=np.float64)
    pi_net = np.empty((L, int(n_grid)), dtype=np.float64) if want_profile else None
    for k, idx in enumerate(parts):
        merged[idx] = pack_small(results[k])
        if pi_net is not None:
            pi_net[idx] = results[k].pi_net
    return unpack_small(merged, K, E, pi_net)


def small_layout(L, K, E):
    """Byte offsets of the nine sections of a batch's small-result block (tapir_cuda_device_small_block) for L loci:
    every section starts on a 256-byte boundary, in the order of Engine.SMALL_SECTIONS."""
    sizes = [L * K * 8, L * E * 8, L * E * 8, L * 8, L * 8, L * 8, L * 8, L * 8, L * 4]
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += (s + 255) & ~255
    return offs, o


_SECTION_DTYPE = {"pi_times": np.float64, "integral": np.float64, "abserr": np.float64, "max_pi": np.float64,
                  "argmax_t": np.int64, "mean_rate": np.float64, "n_finite": np.int64, "n_multi_panel": np.int64,
                  "argmax_epoch": np.int32}


def split_small_block(raw, L, K, E):
    """The nine per-locus vectors held in one small-result block (uint8 array) -> dict of numpy views."""
    from .engine import Engine
    offs, _ = small_layout(L, K, E)
    width = {"pi_times": K, "integral": E, "abserr": E}
    out = {}
    for name, o in zip(Engine.SMALL_SECTIONS, offs):
        dt = np.dtype(_SECTION_DTYPE[name])
        w = width.get(name, 1)
        v = raw[o:o + L * w * dt.itemsize].view(dt)
        out[name] = v.reshape(L, w) if name in width else v
    return out


class ShardedBatch:
    """ONE CSR batch split over the GPUs of a torch.distributed group (one process per GPU) -- the device form of
    Pool.map over loci (bin/tapir_compute.py:159-164).  Rank r keeps partition r (partition_loci: size-balanced, LPT)
    resident on its GPU; the per-locus result vectors of every rank land in one buffer on rank 0's GPU, pushed over
    NVLink by the producing GPU at the end of its kernels (CUDA IPC peer memory), or by one NCCL gather per step when
    peer memory is unavailable; collect() on rank 0 returns them in the ORIGINAL locus order (db ids follow that order,
    tapir/db.py:47-49).  The profile stays where it was computed (each rank copies its rows to its own host).
    Without an initialised process group it is the single-GPU case (one part)."""

    def __init__(self, eng, offsets, rates, n_grid, times, epochs, gather="push", want_profile=True):
        import torch.distributed as dist
        self.eng, self.dist = eng, dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        offsets = np.asarray(offsets, dtype=np.int64)
        self.L, self.T, self.K, self.E = offsets.size - 1, int(n_grid), len(times), len(epochs)
        self.times, self.epochs, self.want_profile = list(times), [list(e) for e in epochs], want_profile
        sizes = offsets[1:] - offsets[:-1]
        self.parts = partition_loci(sizes, self.world)
        self.mine = self.parts[self.rank]
        self.loads = np.array([int(sizes[p].sum()) for p in self.parts])
        self.offsets, self.rates = shard_csr(offsets, rates, self.mine, pinned=True)
        eng.upload(self.offsets, self.rates, self.T, self.times, self.epochs, want_profile=want_profile)
        self.slot_bytes = max(small_layout(len(p), self.K, self.E)[1] for p in self.parts)
        self.mode = "none"
        if self.dist is not None:
            self.mode = self._setup_gather(gather)

    def _setup_gather(self, gather):
        import torch
        dist, eng = self.dist, self.eng
        if gather == "push":
            try:
                obj = [eng.gather_create(self.slot_bytes, self.world) if self.rank == 0 else None]
                dist.broadcast_object_list(obj, src=0)
                eng.gather_attach(None if self.rank == 0 else obj[0], self.slot_bytes, self.rank)
                ok = torch.tensor([1], device="cuda")
            except Exception as err:                      # noqa: BLE001 - any failure means: use NCCL
                import sys
                sys.stderr.write("rank %d: push gather unavailable (%s); falling back to NCCL\n" % (self.rank, err))
                ok = torch.tensor([0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                return "push"
            eng.gather_detach()
        self._nccl = {}
        return "nccl"

    # ---- one step with the batch resident: kernels (+ push) on this rank's GPU ----
    def step(self):
        return self.eng.compute()

    def _nccl_gather(self):
        """One dist.gather of the device block (padded to the largest part) to rank 0."""
        import torch
        ptr, nbytes, _ = self.eng.device_small_block()
        st = self._nccl
        if st.get("key") != (ptr, nbytes):
            class _Dev:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
            st["key"] = (ptr, nbytes)
            st["src"] = torch.as_tensor(_Dev(), device="cuda")
            st["pad"] = torch.zeros(self.slot_bytes, dtype=torch.uint8, device="cuda")
            st["dst"] = [torch.empty_like(st["pad"]) for _ in range(self.world)] if self.rank == 0 else None
        st["pad"][:nbytes].copy_(st["src"])
        self.dist.gather(st["pad"], st["dst"], dst=0)
        return st["dst"]

    def gather_step(self):
        """NCCL mode only: the collective that follows a step (push mode needs none)."""
        if self.mode == "nccl":
            self._nccl_gather()

    # ---- the whole path with host buffers for this rank's partition ----
    def run_host(self, out=None, pinned_results=True):
        return self.eng.run(self.offsets, self.rates, self.T, self.times, self.epochs, want_profile=self.want_profile,
                            out=out, pinned_results=pinned_results)

    # ---- rank 0: every locus' small vectors, original order ----
    def collect(self, local=None):
        """After the ranks have finished a step (barrier inside): rank 0 returns a dict of the nine per-locus vectors
        of the WHOLE batch in the original locus order; other ranks return None.  `local`: this rank's BatchResult,
        used instead of the gather buffer in the single-process case."""
        if self.dist is None:
            res = local if local is not None else self.eng.download()
            return {name: np.asarray(getattr(res, name)) for name in self.eng.SMALL_SECTIONS}
        if self.mode == "push":
            import torch
            torch.cuda.synchronize()
            self.dist.barrier()
            if self.rank != 0:
                return None
            raw = self.eng.gather_read(self.slot_bytes * self.world)
            blocks = [raw[r * self.slot_bytes:(r + 1) * self.slot_bytes] for r in range(self.world)]
        else:
            dst = self._nccl_gather()
            if self.rank != 0:
                return None
            blocks = [d.cpu().numpy() for d in dst]
        full = None
        for r, idx in enumerate(self.parts):               # un-permute to the original locus order
            part = split_small_block(blocks[r], len(idx), self.K, self.E)
            if full is None:
                full = {n: np.empty((self.L,) + v.shape[1:], dtype=v.dtype) for n, v in part.items()}
            for n, v in part.items():
                full[n][idx] = v
        return full

    def close(self):
        """Release the gather buffer: the importing ranks unmap it before its owner frees it."""
        if self.dist is None or self.mode != "push":
            return
        import torch
        torch.cuda.synchronize()
        self.dist.barrier()
        if self.rank != 0:
            self.eng.gather_detach()
        self.dist.barrier()
        if self.rank == 0:
            self.eng.gather_detach()
        self.mode = "closed"
