"""World-size-2 gloo test (CPU) of the locus-sharded driver: partition, per-rank CSR slices, padded
gather of the small result vectors to rank 0 and the un-permutation back to the original locus order.
The per-rank arithmetic is the CPU oracle here (there is no GPU in this container); on GPUs the same
code runs with Engine.run and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_compute(off, rates, T, times, epochs):
    from oracle import capi
    from tapir_b200.engine import BatchResult
    L = off.size - 1
    pi_net, si, se = capi.batch(off, rates, T, epochs)
    fin = np.array([np.isfinite(rates[off[l]:off[l + 1]]).sum() for l in range(L)], dtype=np.int64)
    with np.errstate(invalid="ignore"):
        mean = np.array([np.nanmean(rates[off[l]:off[l + 1]]) if off[l + 1] > off[l] else np.nan for l in range(L)])
    return BatchResult(pi_net=pi_net, pi_times=pi_net[:, times], integral=si, abserr=se, max_pi=pi_net.max(axis=1),
                       argmax_t=pi_net.argmax(axis=1).astype(np.int64), argmax_epoch=si.argmax(axis=1).astype(np.int32),
                       mean_rate=mean, n_finite=fin, n_multi_panel=np.zeros(L, dtype=np.int64))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from tapir_b200 import sharded, synthetic
    from tapir_b200.engine import pack_csr
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        loci = [synthetic.uce_rates(rng, int(n), 120, nan_frac=0.05) for n in rng.integers(5, 90, size=23)]
        off, rates = pack_csr(loci, pinned=False)
        times, epochs = [3, 100], [[32, 42], [88, 98], [95, 105]]
        res = sharded.run_sharded(_oracle_compute, off, rates, 120, times, epochs, gather_profile=True)
        if rank == 0:
            ref = _oracle_compute(off, rates, 120, times, epochs)
            ok = all(np.array_equal(getattr(res, f), getattr(ref, f), equal_nan=True) if getattr(ref, f).dtype.kind == "f"
                     else np.array_equal(getattr(res, f), getattr(ref, f))
                     for f in sharded.SMALL_FIELDS) and np.array_equal(res.pi_net, ref.pi_net)
            with open(out_path, "w") as fh:
                fh.write("ok" if ok else "mismatch")
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_csr_and_pack_roundtrip():
    from tapir_b200 import sharded
    from tapir_b200.engine import partition_loci
    off = np.array([0, 3, 3, 10, 12], dtype=np.int64)
    rates = np.arange(12, dtype=np.float64)
    parts = partition_loci(off[1:] - off[:-1], 2)
    assert sorted(np.concatenate(parts).tolist()) == [0, 1, 2, 3]
    for idx in parts:
        o, r = sharded.shard_csr(off, rates, idx)
        for j, l in enumerate(idx):
            assert np.array_equal(r[o[j]:o[j + 1]], rates[off[l]:off[l + 1]])
    res = _oracle_compute(off, rates * 1e-3, 20, [1, 5], [[0, 10], [3, 4]])
    back = sharded.unpack_small(sharded.pack_small(res), 2, 2)
    for f in sharded.SMALL_FIELDS:
        assert np.array_equal(getattr(back, f), getattr(res, f), equal_nan=True), f
