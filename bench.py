#!/usr/bin/env python
"""bench.py -- Townsend-PI hot path on B200: site*timepoint PI evals/s and loci/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one pass of the whole hot path (PI profile + --times picks + GK21/QAGS epoch
integrals + per-locus maxima) over one batch of synthetic loci.  At N = 1 the batch is
BASELINE.json configs[1] (5,000 loci x 500 sites, 1,000-point grid, 4 epochs); at N > 1 every
rank owns one such shard (weak scaling; loci are independent, bin/tapir_compute.py:159-164) and
the small per-locus result vectors are gathered to rank 0 over NCCL.

Printed keys (one JSON line, rank 0):
  value    evals/s with inputs resident in HBM, kernels timed by CUDA events on the library's stream
  e2e      the same metric through the C-ABI call with HOST buffers: pinned H2D + kernels + D2H
  roofline FP64-pipe roofline of the dominant kernel (profile_kernel), see DESIGN.md
  cpu_baseline  the reference's CPU path (numpy + scipy quad port, Pool(cpu_count-1)) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "site*timepoint PI evals/sec"
UNIT = "evals/s"


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm (oracle port: numpy ufuncs + scipy.integrate.quad),
# mapped over loci with multiprocessing.Pool(cpu_count()-1) as bin/tapir_compute.py:162-164 does.
# --------------------------------------------------------------------------------------------
_CPU_STATE = {}


def _cpu_init(T, times, epochs):
    from oracle import townsend_port as tp
    _CPU_STATE["tp"] = tp
    _CPU_STATE["tcol"] = tp.time_column(0, T)
    _CPU_STATE["times"] = times
    _CPU_STATE["epochs"] = epochs


def _cpu_worker(lam):
    import warnings
    tp = _CPU_STATE["tp"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # returns the full (T,S) matrix through the pool pipe, as the reference's worker does
        return tp.locus_worker(_CPU_STATE["tcol"], lam, _CPU_STATE["times"], _CPU_STATE["epochs"], keep_matrix=True)


def cpu_reference_pass(offsets, rates, T, times, epochs, n_loci, pool):
    loci = [rates[offsets[l]:offsets[l + 1]] for l in range(n_loci)]
    t0 = time.perf_counter()
    out = pool.map(_cpu_worker, loci)
    dt = time.perf_counter() - t0
    evals = int(offsets[n_loci] - offsets[0]) * T
    return dt, evals, out


def make_pool(T, times, epochs):
    import multiprocessing as mp
    workers = max(1, (os.cpu_count() or 2) - 1)
    ctx = mp.get_context("fork")
    return ctx.Pool(processes=workers, initializer=_cpu_init, initargs=(T, times, epochs)), workers


def cpu_baseline(offsets, rates, T, times, epochs, budget_s=15.0):
    """Bounded sample: calibrate on one pool round, then size the sample for ~budget_s seconds."""
    pool, workers = make_pool(T, times, epochs)
    try:
        L = offsets.size - 1
        n0 = min(L, 4 * workers)
        cpu_reference_pass(offsets, rates, T, times, epochs, min(L, workers), pool)      # fork + import warm-up
        dt0, _, _ = cpu_reference_pass(offsets, rates, T, times, epochs, n0, pool)
        n = int(min(L, max(n0, n0 * budget_s / max(dt0, 1e-3))))
        dt, evals, _ = cpu_reference_pass(offsets, rates, T, times, epochs, n, pool)
    finally:
        pool.close()
        pool.join()
    return {
        "value": evals / dt, "unit": UNIT, "cores": workers, "kind": "port",
        "loci_per_s": n / dt,
        "sample": "first %d of %d loci of the same batch, %.1f s wall, Pool(%d) over numpy+scipy.quad port "
                  "(oracle/townsend_port.py), full (T,S) matrix returned through the pool pipe as the reference does"
                  % (n, L, dt, workers),
    }


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every few ms on a host thread while the timed
    region runs (a step is well under a millisecond, far below nvidia-smi's polling period)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0, period_s=0.002):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.mask, self.smax = [], 0, None
        self._stop = threading.Event()
        self.thread = None
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.idx
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[self.idx])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self.nv is None or self.thread is None:
            return None
        self._stop.set()
        self.thread.join(timeout=2)
        if not self.sm:
            return None
        hi = max(self.sm)
        load = [x for x in self.sm if x >= 0.6 * hi]    # "under load": within 60 % of the busiest clock seen
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": self.smax,
                "reasons": sorted(k for k, bit in self.REASONS.items() if self.mask & bit),
                "samples": len(self.sm), "source": "NVML, %.0f ms period, sampled over the attribution, timed and end-to-end loops" % (self.period * 1e3)}


def profile_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one profile-kernel launch from the committed ncu --set full
    capture of this workload (profiles/*_ncu_full_<workload>.txt), or None."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_%s.txt" % workload)))
    if not files:
        return None
    total, inside = 0.0, False
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(files[-1]):
        if line.startswith("== "):
            inside = "profile_kernel" in line or "profile_mom_kernel" in line
        m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if inside and m:
            total += float(m.group(2)) * scale.get(m.group(3), 1.0)
    return total or None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the CPU path alone, rank 0 only."""
    if rank != 0:
        return
    from tapir_b200 import synthetic
    off, rates, T = synthetic.build(args.workload)
    times, epochs = synthetic.TIMES, synthetic.EPOCHS
    pool, workers = make_pool(T, times, epochs)
    L = off.size - 1
    try:
        # size one step for ~8 s so that steps + warmup finish within a few minutes
        cpu_reference_pass(off, rates, T, times, epochs, min(L, workers), pool)           # fork + import warm-up
        n0 = min(L, 4 * workers)
        dt0, _, _ = cpu_reference_pass(off, rates, T, times, epochs, n0, pool)
        n = int(min(L, max(n0, n0 * 8.0 / max(dt0, 1e-3))))
        for _ in range(args.warmup):
            cpu_reference_pass(off, rates, T, times, epochs, min(n, workers), pool)
        tot_t = tot_e = 0.0
        for _ in range(args.steps):
            dt, evals, _ = cpu_reference_pass(off, rates, T, times, epochs, n, pool)
            tot_t += dt
            tot_e += evals
    finally:
        pool.close()
        pool.join()
    value = tot_e / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "sample_loci_per_step": n, "timing": "host wall clock"},
        "loci_per_s": n * args.steps / tot_t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": "%d of %d loci per step, Pool(%d), numpy+scipy.quad port of tapir/compute.py" % (n, L, workers)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(w):
    from tapir_b200 import synthetic
    c = synthetic.CONFIGS[w]
    if c["kind"] == "uniform":
        return "%s: synthetic %d loci x %d sites, %d-point grid, 4 epochs, 4 times" % (w, c["n_loci"], c["n_sites"], c["T"])
    return "%s: synthetic %d loci x ~%d sites (ragged %d-%d), %d-point grid, 4 epochs, 4 times" % (
        w, c["n_loci"], c["mean_sites"], c["smin"], c["smax"], c["T"])


def host_side_rates(rates_np, off_np, T, times, epochs, out, n_files=512, n_db=256):
    """JSON ingest and the sqlite sink are outside the timed region and reported separately.
    Bounded samples: n_files rate files of this workload's shape parsed by the library's C++ reader (and by
    json.load for comparison); n_db loci of this step's results written with the bulk sqlite writer."""
    import json as _json
    import shutil
    import tempfile
    from tapir_b200 import compute, db, ingest
    from tapir_b200.engine import BatchResult
    d = tempfile.mkdtemp(prefix="tapir_bench_")
    try:
        # bounded by content, not only by count: at most ~1M sites of rate files and ~1M sqlite rows (the genome-scale
        # shape has 1e6 sites per locus and a 10,000-point grid; 512 such files would take minutes to write and parse)
        n_files = min(n_files, off_np.size - 1)
        while n_files > 1 and off_np[n_files] > 1_000_000:
            n_files //= 2
        n_db = max(1, min(n_db, 1_000_000 // max(int(T), 1)))
        paths = []
        for l in range(n_files):
            lam = rates_np[off_np[l]:off_np[l + 1]]
            doc = {"sites": {"rates": [{"site": k + 1, "subst": 0.0, "rate": float(v), "ll": 0.0} for k, v in enumerate(lam)]}}
            p = os.path.join(d, "locus-%d.rates" % l)
            with open(p, "w") as fh:
                _json.dump(doc, fh)
            paths.append(p)
        t0 = time.perf_counter()
        off_i, r_i = ingest.read_rate_files(paths, correction=1, pinned=False)
        t_native = time.perf_counter() - t0
        ok = bool(np.array_equal(r_i, rates_np[:off_np[n_files]], equal_nan=True))
        n_py = min(64, n_files)
        t0 = time.perf_counter()
        for p in paths[:n_py]:
            compute.parse_site_rates(p, 1, test=True)
        t_py = (time.perf_counter() - t0) * n_files / n_py
        n_db = min(n_db, out.integral.shape[0])
        sub = BatchResult(pi_net=np.array(out.pi_net[:n_db]), pi_times=np.array(out.pi_times[:n_db]),
                          integral=np.array(out.integral[:n_db]), abserr=np.array(out.abserr[:n_db]), max_pi=None,
                          argmax_t=None, argmax_epoch=None, mean_rate=None, n_finite=None, n_multi_panel=None)
        conn, c = db.create_probe_db(os.path.join(d, "pi.sqlite"))
        t0 = time.perf_counter()
        db.insert_batch(conn, c, ["locus-%d.rates" % l for l in range(n_db)], sub, times, epochs)
        conn.commit()
        t_db = time.perf_counter() - t0
        conn.close()
        rows = n_db * (1 + T + len(times) + len(epochs))
        return {"ingest_files": n_files, "ingest_native_files_per_s": n_files / t_native, "ingest_native_matches": ok,
                "ingest_json_load_files_per_s": n_files / t_py, "sqlite_loci": n_db, "sqlite_rows_per_s": rows / t_db,
                "sqlite_loci_per_s": n_db / t_db,
                "note": "host side of the drop-in, outside the timed region: tapir_ingest_rates (C++ threads) vs json.load "
                        "per file; bulk executemany writer with the reference's schema incl. the T-rows-per-locus net table"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_ours(args, rank, world, local_rank):
    from tapir_b200 import synthetic
    from tapir_b200 import sharded
    from tapir_b200.engine import Engine, bind_host_thread_near, pinned_empty

    times, epochs = synthetic.TIMES, synthetic.EPOCHS
    # every rank owns one shard of the global batch (same shape, different seed)
    cfg = dict(synthetic.CONFIGS[args.workload])
    cfg["seed"] = cfg["seed"] * 1000 + rank
    synthetic.CONFIGS["_shard"] = cfg
    off_np, rates_np, T = synthetic.build("_shard")
    L, nnz = off_np.size - 1, rates_np.size
    evals_rank = nnz * T
    # sites that contribute: a zero rate (invariant site) or a culled (NaN/Inf) site adds exactly 0 everywhere; the
    # kernels drop them after the load.  `value` counts every site*timepoint (BASELINE.json's metric); the roofline
    # counts only the FMAs that are algorithmically required, i.e. those of the active sites.
    with np.errstate(over="ignore", invalid="ignore"):
        a_np = 16.0 * rates_np * rates_np
    n_active = int(np.count_nonzero(np.isfinite(a_np) & (a_np != 0.0)))
    n_slow = int(np.count_nonzero(np.isfinite(a_np) & (a_np != 0.0) & (rates_np >= 0.0) & (rates_np < 0.01953125)))
    evals_active = n_active * T

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(off_np, rates_np, T, times, epochs)   # before any CUDA call (fork safety)

    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    numa_bound = bind_host_thread_near(local_rank) if world > 1 else False   # after the CPU pool (N = 1 keeps all cores)
    eng = Engine(local_rank)
    offsets = pinned_empty(off_np.shape, np.int64); offsets[:] = off_np
    rates = pinned_empty(rates_np.shape, np.float64); rates[:] = rates_np

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- small-result gather (N > 1).  The library keeps every per-locus vector of a batch in ONE device block.
    #  "push" (default): rank 0 owns a buffer with one slot per rank, exported through CUDA IPC; every rank's
    #      tapir_cuda_compute ends with a device-to-device copy of its block into its slot over NVLink, on the
    #      same stream as finalize -- no collective launch inside a step.
    #  "nccl": one dist.gather of the block per step (fallback if IPC peer memory is unavailable, or --gather nccl)
    gather = None
    gather_mode = "none"
    if dist is not None:
        eng.upload(offsets, rates, T, times, epochs, want_profile=True)
        _, slot_bytes, sections = eng.device_small_block()
        gather_mode = args.gather
        if gather_mode == "push":
            try:
                obj = [eng.gather_create(slot_bytes, world) if rank == 0 else None]
                dist.broadcast_object_list(obj, src=0)
                eng.gather_attach(None if rank == 0 else obj[0], slot_bytes, rank)
                ok = torch.tensor([1], device="cuda")
            except Exception as err:                      # noqa: BLE001 - any failure means: use NCCL
                sys.stderr.write("rank %d: push gather unavailable (%s); falling back to NCCL\n" % (rank, err))
                ok = torch.tensor([0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                eng.gather_detach()
                gather_mode = "nccl"
        if gather_mode == "nccl":
            class _Dev:
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

            state = {}

            def gather():
                ptr, nbytes, _ = eng.device_small_block()
                if state.get("ptr") != (ptr, nbytes):
                    state["ptr"] = (ptr, nbytes)
                    state["src"] = torch.as_tensor(_Dev(ptr, nbytes // 8), device="cuda")
                    state["dst"] = [torch.empty_like(state["src"]) for _ in range(world)] if rank == 0 else None
                dist.gather(state["src"], state["dst"], dst=0)
                return state["dst"]

    # ---- device-resident timing ----
    eng.upload(offsets, rates, T, times, epochs, want_profile=True)
    # clocks and throttle reasons are sampled (NVML, 1 ms period) from here to the end of the end-to-end loop: the
    # attribution pass, the timed steps and the e2e steps all run the same kernels under load
    sampler = ClockSampler(local_rank, period_s=0.001)
    if rank == 0:
        sampler.start()
    # (1) attribution pass: kernels serialised on one stream so that per-kernel CUDA-event times are exact
    eng.set_serial_kernels(True)
    for _ in range(max(args.warmup, 3)):
        eng.compute()
    prof_ms = quad_ms = fin_ms = serial_ms = 0.0
    for _ in range(args.steps):
        eng.flush_l2()
        tm = eng.compute()
        prof_ms += tm.profile_ms; quad_ms += tm.quad_ms; fin_ms += tm.finalize_ms; serial_ms += tm.total_ms
    # (2) the timed steps proper: the shipped configuration (profile and integral kernels on two streams)
    eng.set_serial_kernels(False)
    for _ in range(max(args.warmup, 3)):
        eng.compute()
        if gather:
            gather()
    barrier()
    dev_ms, gather_ms, launches = 0.0, 0.0, 0
    for _ in range(args.steps):
        eng.flush_l2()                      # untimed: the step starts with a cold L2
        tm = eng.compute()                  # CUDA events on the library stream, synchronised on both sides
        dev_ms += tm.total_ms
        launches += tm.launches
        if gather:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gather(); e1.record(); torch.cuda.synchronize()
            gather_ms += e0.elapsed_time(e1)
    barrier()
    step_ms = (dev_ms + gather_ms) / args.steps
    gather_check = None
    if dist is not None:
        # every rank's integral checksum, to compare with what arrived at rank 0
        mine = eng.download()
        sums = [None] * world
        dist.all_gather_object(sums, float(mine.integral.sum()))
        if rank == 0:
            E = len(epochs)
            if gather_mode == "push":
                raw = eng.gather_read(slot_bytes * world)
                blocks = [raw[r * slot_bytes:(r + 1) * slot_bytes] for r in range(world)]
            else:
                blocks = [d.cpu().numpy().view(np.uint8) for d in gather()]
            got = [float(b[sections["integral"]:sections["integral"] + L * E * 8].view(np.float64).sum()) for b in blocks]
            gather_check = "ok" if np.allclose(got, sums, rtol=1e-12) else "MISMATCH %r vs %r" % (got, sums)
        elif gather_mode == "nccl":
            gather()

    # ---- end to end through the public call: pinned host in, pinned host out ----
    out = eng.run(offsets, rates, T, times, epochs, want_profile=True, pinned_results=True)   # allocates + warms
    for _ in range(2):
        eng.run(offsets, rates, T, times, epochs, want_profile=True, out=out)
    h0, d0 = eng.transfer_bytes()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run(offsets, rates, T, times, epochs, want_profile=True, out=out)
        if gather_mode == "nccl":
            # the step's results are on the host already (tapir_cuda_run); the per-locus vectors go through NCCL
            packed = torch.from_numpy(sharded.pack_small(out)).cuda()
            dst = [torch.empty_like(packed) for _ in range(world)] if rank == 0 else None
            dist.gather(packed, dst, dst=0)
            if rank == 0:
                _ = [d.cpu() for d in dst]
            torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    h1, d1 = eng.transfer_bytes()
    clocks = sampler.stop() if rank == 0 else None
    # PCIe context for the e2e number: the D2H copy of one step's results on its own (pinned destination)
    eng.upload(offsets, rates, T, times, epochs, want_profile=True)
    eng.compute()
    eng.download(out=out)
    _, dd0 = eng.transfer_bytes()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.download(out=out)
    d2h_s = (time.perf_counter() - t0) / 5
    _, dd1 = eng.transfer_bytes()
    d2h_gbs = (dd1 - dd0) / 5 / d2h_s / 1e9
    checksum = float(out.integral.sum())

    peak_tf, _ = eng.fp64_peak()

    if dist is not None:
        t = torch.tensor([step_ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, e2e_s = float(t[0]), float(t[1])
        dist.destroy_process_group()
    if rank != 0:
        return

    total_evals = evals_rank * world
    total_loci = L * world
    peaks, peak_src = measured_peaks()
    prof_s = prof_ms / args.steps * 1e-3
    # what tapir_cuda_compute reports it launched (tapir_timing), not a guess from the grid size
    kernel_name = "%s<%d>" % ("profile_mom_kernel" if tm.profile_moment_form else "profile_kernel", tm.profile_tile_points // 256)
    achieved_tf = 2.0 * evals_active / prof_s / 1e12      # one FP64 FMA per (active site, timepoint): DESIGN.md
    line = {
        "metric": METRIC, "value": total_evals / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload) + (" per GPU" if world > 1 else ""),
                   "loci_per_gpu": L, "grid": T, "epochs": epochs, "times": times,
                   "cache": "L2 flushed (256 MB memset) before every timed step",
                   "timing": "CUDA events on the launching stream per step, max over ranks",
                   "parallelism": ("locus-sharded x%d, per-locus vectors gathered to rank 0 by %s" % (
                       world, "NVLink peer-memory push fused after finalize (CUDA IPC)" if gather_mode == "push"
                       else "one NCCL gather per step")) if world > 1 else "single GPU"},
        "loci_per_s": total_loci / (step_ms * 1e-3),
        # the same step counted over the sites that need arithmetic only (non-zero finite rate): `value` follows the
        # metric's definition (every site*timepoint of the input), this one is what the FP64 pipe actually evaluates
        "active_evals_per_s": evals_active * world / (step_ms * 1e-3),
        "kernel_ms": {"profile": prof_ms / args.steps, "quad": quad_ms / args.steps, "finalize": fin_ms / args.steps,
                      "serialised_step": serial_ms / args.steps, "gather": gather_ms / args.steps,
                      "note": "per-kernel times from a separate pass with the kernels serialised on one stream; "
                              "ms_per_step is the shipped two-stream configuration"},
        "roofline": {
            "bound": "tensor", "pipe": "FP64 (DMMA sub-pipe of the tensor pipe; shared with DFMA on B200)",
            "kernel": kernel_name, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": profile_traffic(args.workload),
            "peak_source": "FP64 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has HBM and bf16 entries only; "
                           "tools/microbench/fp64_pipes.cu: DMMA 37.0, DFMA 36.6 TFLOP/s, one shared pipe)",
            "algorithmic": "2 flop (1 FP64 FMA) per site*timepoint eval of a site with a non-zero finite rate: "
                           "%d of the launch's %d evals (zero-rate and culled sites add exactly 0 and are dropped after "
                           "the load).  profile_mom_kernel EXECUTES about a quarter of them: a site with rate < 0.0195 "
                           "enters the contraction as 32 Taylor moments per 128-point block of the time tile instead of 32 "
                           "powers per 32-point block (DESIGN.md 3.1), so this fraction measures the direct formulation's "
                           "work against the FP64 peak and may exceed 1" % (evals_active, evals_rank),
            "executed_over_algorithmic": None if T < 1000 else round(
                (0.25 * n_slow + (n_active - n_slow)) / max(n_active, 1), 4),
            "active_site_fraction": n_active / max(nnz, 1),
            "quad_kernel": {"ms": quad_ms / args.steps, "site_epochs_per_s": nnz * len(epochs) / (quad_ms / args.steps * 1e-3)},
            # the same evals/s against the ceiling of the reference's formulation, one FP64 exp per
            # (site, timepoint): ~17 pipe operations for a branch-free exp + 3 for a*t*e, on the same measured pipe
            "naive_one_exp_per_eval": {"ops_per_eval": 20, "ceiling_evals_per_s": peak_tf * 1e12 / 2.0 / 20.0,
                                       "value_over_ceiling": (evals_rank / (step_ms * 1e-3)) / (peak_tf * 1e12 / 2.0 / 20.0)
                                       if peak_tf else None},
            "hbm": {"achieved_gbs": 8.0 * nnz / prof_s / 1e9, "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src,
                    "note": "rate stream, 8 B per site per time tile"},
        },
        "e2e": {"value": total_evals / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": (h1 - h0) // args.steps, "d2h_bytes_per_step": (d1 - d0) // args.steps,
                "loci_per_s": total_loci / e2e_s, "timing": "host wall clock around tapir_cuda_run, pinned buffers",
                "d2h_alone_ms": d2h_s * 1e3, "d2h_gbs": d2h_gbs, "host_thread_bound_to_gpu_numa_node": numa_bound,
                "note": "chunks of loci (13 MB of traffic each, a quarter- and a half-size one first) rotate over up to eight lanes (H2D / kernels / D2H of consecutive chunks overlap); bounded below by the D2H copy of the profile"},
        "gpu_launches": launches, "gather": {"mode": gather_mode, "check": gather_check},
        "clocks": clocks,
        "checksum": checksum,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_host_side:
        line["host_side"] = host_side_rates(rates_np, off_np, T, times, epochs, out)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-side", action="store_true", help="skip the (untimed) ingest / sqlite sample")
    ap.add_argument("--gather", default="push", choices=["push", "nccl"], help="N > 1: how per-locus vectors reach rank 0")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
