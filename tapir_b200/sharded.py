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
