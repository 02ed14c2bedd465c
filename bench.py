#!/usr/bin/env python
"""bench.py -- Townsend-PI hot path on B200: site*timepoint PI evals/s and loci/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one pass of the whole hot path (PI profile + --times picks + QUADPACK epoch integrals + per-locus
maxima) over one batch of synthetic loci.  The batch is BASELINE.json configs[4] -- 500,000 loci x 800 sites on a
2,048-point grid, the configuration the metric ("... at 1/2/4/8 B200") is quoted on -- at every N: ONE batch, split
over the ranks by the size-balanced locus partition (strong scaling; loci are independent,
bin/tapir_compute.py:159-164), per-locus result vectors gathered to rank 0 over NVLink and put back into the original
locus order there.  configs[1] and configs[2] ride along as `other_configs` (N = 1); at N > 1 configs[2] goes through
the LPT partition and configs[1] runs as one shard per rank (the round-1 weak-scaling number).

Printed keys (one JSON line, rank 0):
  value         evals/s with inputs resident in HBM, kernels timed by CUDA events on the library's stream, max over ranks
  e2e           the same metric through the C-ABI call with HOST buffers: pinned H2D + kernels + D2H (+ gather, un-permute)
  roofline      FP64-pipe roofline of the dominant kernel: executed lane-operations, and the direct formulation's beside it
  parity        sampled loci of the timed batch against the CPU oracle, outside the timed region
  cpu_baseline  the reference's CPU path (Pool(cpu_count-1)) on a bounded sample of the same batch
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "site*timepoint PI evals/sec"
UNIT = "evals/s"
DEFAULT_WORKLOAD = "cfg5"


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm mapped over loci with multiprocessing.Pool(cpu_count()-1) as
# bin/tapir_compute.py:162-164 does.  When build() has placed the reference's unmodified tapir/compute.py under
# baseline/_ref/ (git-ignored; it travels to the GPU box) the arithmetic is THAT file's functions, imported with the
# three-line shim of SURVEY.md 8c (kind "reference"); otherwise the bit-exact port oracle/townsend_port.py (kind "port").
# --------------------------------------------------------------------------------------------
_CPU_STATE = {}
REF_COMPUTE = os.path.join(ROOT, "baseline", "_ref", "tapir", "compute.py")


def load_reference_compute():
    """The reference's tapir/compute.py, unmodified, or None."""
    if not os.path.exists(REF_COMPUTE):
        return None
    import importlib.util
    import types
    import numpy
    import scipy
    sys.modules.setdefault("dendropy", types.ModuleType("dendropy"))   # only correct_branch_lengths needs it
    if not hasattr(scipy, "vectorize"):
        scipy.vectorize = numpy.vectorize                               # removed from modern scipy (compute.py:20)
    spec = importlib.util.spec_from_file_location("tapir_ref_compute", REF_COMPUTE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_kind():
    return "reference" if os.path.exists(REF_COMPUTE) else "port"


def _cpu_init(T, times, epochs):
    ref = load_reference_compute()
    _CPU_STATE["ref"] = ref
    if ref is None:
        from oracle import townsend_port as tp
        _CPU_STATE["tp"] = tp
        _CPU_STATE["tcol"] = tp.time_column(0, T)
    else:
        _CPU_STATE["tcol"] = ref.get_time(0, T)
    _CPU_STATE["times"] = times
    _CPU_STATE["epochs"] = epochs


def _cpu_worker(lam):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = _CPU_STATE["ref"]
        tcol, times, epochs = _CPU_STATE["tcol"], _CPU_STATE["times"], _CPU_STATE["epochs"]
        if ref is None:
            # returns the full (T,S) matrix through the pool pipe, as the reference's worker does
            return _CPU_STATE["tp"].locus_worker(tcol, lam, times, epochs, keep_matrix=True)
        # worker() of bin/tapir_compute.py:110-123 on the reference's own functions
        mean_rate = np.mean(np.ma.masked_array(lam, np.isnan(lam)))
        pi = ref.get_townsend_pi(tcol, lam)
        pi_net = np.nansum(pi, axis=1)
        pi_times = ref.get_net_pi_for_periods(pi, times)
        pi_epochs = ref.get_net_integral_for_epochs(lam[np.isfinite(lam)], epochs)
        return lam, mean_rate, pi, pi_net, pi_times, pi_epochs


def cpu_reference_pass(offsets, rates, T, n_loci, pool):
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


def cpu_how():
    return ("the reference's unmodified tapir/compute.py (baseline/_ref, dendropy stubbed)" if cpu_kind() == "reference"
            else "numpy+scipy.quad port of tapir/compute.py (oracle/townsend_port.py)")


def cpu_baseline(offsets, rates, T, times, epochs, budget_s=15.0):
    """Bounded sample: calibrate on one pool round, then size the sample for ~budget_s seconds."""
    pool, workers = make_pool(T, times, epochs)
    try:
        L = offsets.size - 1
        n0 = min(L, 4 * workers)
        cpu_reference_pass(offsets, rates, T, min(L, workers), pool)      # fork + import warm-up
        dt0, _, _ = cpu_reference_pass(offsets, rates, T, n0, pool)
        n = int(min(L, max(n0, n0 * budget_s / max(dt0, 1e-3))))
        dt, evals, _ = cpu_reference_pass(offsets, rates, T, n, pool)
    finally:
        pool.close()
        pool.join()
    return {"value": evals / dt, "unit": UNIT, "cores": workers, "kind": cpu_kind(), "loci_per_s": n / dt,
            "sample": "first %d of %d loci of the same batch, %.1f s wall, Pool(%d) over %s, full (T,S) matrix returned "
                      "through the pool pipe as the reference does" % (n, L, dt, workers, cpu_how())}


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every few ms on a host thread while the timed
    region runs (a step is milliseconds, far below nvidia-smi's polling period)."""
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
                "samples": len(self.sm),
                "source": "NVML, %.0f ms period, sampled over the attribution, timed and end-to-end loops" % (self.period * 1e3)}


def profile_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one profile-kernel launch from the committed ncu --set full
    capture of this workload (profiles/*_ncu_full_<workload>.txt, latest round), or None."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_%s.txt" % workload)))
    if not files:
        return None
    total, inside, seen = 0.0, False, 0
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for line in open(files[-1]):
        if line.startswith("== "):
            inside = ("profile_kernel" in line or "profile_mom_kernel" in line) and seen == 0
            seen += 1 if inside else 0
        m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if inside and m:
            total += float(m.group(2)) * scale.get(m.group(3), 1.0)
    return total or None


def profile_ncu_utilisation(workload):
    """Unit utilisations of the profile kernel from the same committed capture (percent of peak): what actually bounds
    the kernel next to the FP64-pipe fraction the roofline object reports.  {} when no capture is committed."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_%s.txt" % workload)))
    if not files:
        return {}
    want = {"sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_subpipe_pct",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "other_fp64_pct",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_pct",
            "sm__icc_request_hit_rate.pct": "instruction_cache_hit_pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct"}
    out, inside, seen = {}, False, 0
    for line in open(files[-1]):
        if line.startswith("== "):
            inside = ("profile_kernel" in line or "profile_mom_kernel" in line) and seen == 0
            seen += 1 if inside else 0
        m = re.match(r"(\S+)\s+([0-9.]+)\s+%", line)
        if inside and m and m.group(1) in want:
            out[want[m.group(1)]] = round(float(m.group(2)), 1)
    if out:
        out["source"] = "profiles/" + os.path.basename(files[-1])
    return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------------------------
# What the profile kernel EXECUTES on the FP64 pipe (profile_mom_kernel / profile_kernel, kernels.cuh), counted from
# the batch itself: DMMA.8x8x4 instructions (256 FMA lanes each) and scalar FP64 warp instructions (32 lanes each).
# The DMMA count follows the kernels' loop structure exactly; the scalar counts per group / per item are read off the
# kernel source.  tests/test_bench_model.py checks both against the opcode counts of the committed ncu capture.
# --------------------------------------------------------------------------------------------
MOM_BOUND = {32: 0.01953125, 16: 0.0048828125}     # moment classes of profile_mom_kernel<RH, KM> (mom_epoch.h)


def site_classes(offsets, rates, moment_form, chunk=1 << 26, coarse=False, tile_points=0):
    """Per locus: (# sites of profile_mom_kernel's moment class, # other sites with a non-zero finite rate).
    moment_form: 0 (direct form: no moment class) or the number of moments per block, 32 or 16 (True = 32).
    coarse=True (2,048-point tiles): a third array, the sites below half the bound (expanded on 256-point blocks),
    which are then not counted in the first.  tile_points > 0: a last array, the whole-tile class (rate below
    0.0048828125 x 128 / tile_points), counted in none of the others."""
    bound = MOM_BOUND[16 if moment_form == 16 else 32]
    off = np.asarray(offsets, dtype=np.int64)
    L = off.size - 1
    n_slow, n_fast, n_coarse = np.zeros(L, dtype=np.int64), np.zeros(L, dtype=np.int64), np.zeros(L, dtype=np.int64)
    n_whole = np.zeros(L, dtype=np.int64)

    def result():
        out = (n_slow, n_fast) + ((n_coarse,) if coarse else ())
        return out + ((n_whole,) if tile_points else ())
    if L == 0 or rates.size == 0:
        return result()
    l0 = 0
    while l0 < L:                                           # whole loci per chunk of at most ~`chunk` sites
        l1 = int(np.searchsorted(off, off[l0] + chunk, side="right")) - 1
        l1 = min(L, max(l1, l0 + 1))
        r = np.asarray(rates[off[l0]:off[l1]])
        mag = np.abs(r)
        active = (mag > 1.5717277847026288e-162) & (mag <= 3.3519519824856493e+153)   # 16 r^2 neither 0 nor Inf (rate_is_active)
        slow = active & (r > 0.0) & (r < bound) if moment_form else np.zeros_like(active)
        seg = off[l0:l1] - off[l0]
        lens = off[l0 + 1:l1 + 1] - off[l0:l1]
        full = np.nonzero(lens > 0)[0]                       # reduceat over the non-empty loci only (their segments tile r)
        whole = slow & (r < 0.0048828125 * 128.0 / tile_points) if tile_points else np.zeros_like(active)
        low = slow & ~whole & (r < 0.5 * bound) if coarse else np.zeros_like(active)
        for mask, dst in ((slow & ~low & ~whole, n_slow), (active & ~slow, n_fast), (low, n_coarse), (whole, n_whole)):
            if full.size:
                dst[l0 + full] = np.add.reduceat(mask, seg[full], dtype=np.int64)
        l0 = l1
    return result()


def executed_fp64_ops(offsets, rates, T, tile_points, moment_form, n_mom_epochs=0):
    """-> dict(dmma, scalar, lane_ops, n_active) of one profile launch over the batch."""
    TT = int(tile_points)
    n_tiles = max(1, (int(T) + TT - 1) // TT)
    RH, NB = TT // 256, TT // 128
    NT = max(1, NB // 8)
    KM = 16 if moment_form == 16 else 32
    two_level = bool(moment_form) and TT == 2048        # MomCfg::Coarse
    n_coarse = n_whole = None
    if two_level:
        n_slow, n_fast, n_coarse, n_whole = site_classes(offsets, rates, moment_form, coarse=True, tile_points=TT)
    elif moment_form:
        n_slow, n_fast, n_whole = site_classes(offsets, rates, moment_form, tile_points=TT)
    else:
        n_slow, n_fast = site_classes(offsets, rates, moment_form)

    def ksteps(cnt, group):
        """k-steps (4 sites each, issued in pairs) summed over the groups of `group` sites of every item."""
        full, last = cnt // group, cnt % group
        ks_last = 2 * ((((last + 3) // 4) + 1) // 2)
        return full * (group // 4) + np.where(last > 0, ks_last, 0)

    if moment_form:
        groups, fgroups = (n_slow + 31) // 32, (n_fast + 15) // 16
        dmma = (KM // 8) * NT * ksteps(n_slow, 32) + 4 * KM * NT + (KM // 4) * NT * ((n_mom_epochs + 7) // 8) + 4 * RH * ksteps(n_fast, 16)
        # scalar FP64 warp instructions (read off kernels.cuh).  Per group of the moment pass: c, a, z (4), the bounded
        # exponential in Estrin form (21), q^2, q^4, z^2, z^4, three moment seeds, KM - 4 chain multiplies, q^8 and the second
        # chain's seed, NB column multiplies (+ a full exponential and a multiply for tiles after the first).  Per item:
        # two accumulator sets added (2 MI NT), 1/k! onto the moments (KS NT), per 1,024-point sub-tile 16 (KS - 1)
        # multiplies of the moments -> times powers (seeds from a table), epilogue 16 per 256 points (four multiplies
        # and four additions per 128-point block).  Per group of 16 fast sites: one / two exponentials + the tables.
        MI, KS, NH = KM // 8, KM // 4, max(1, RH // 4)
        per_group = 30 + KM + NB
        per_item = 2 * MI * NT + KS * NT + NH * 16 * (KS - 1) + 16 * RH
        scalar0 = groups * per_group + per_item + fgroups * NH * 103
        scalar1 = groups * (per_group + 27) + per_item + fgroups * NH * 130
        # the whole-tile class: per group of 32 sites c, a, z, the exponential (21), 16 additions and 17 multiplies, no
        # DMMA; per item with such sites the halving reduction (16 additions), 16 scalings, six DMMAs of the shift and
        # four additions into the moment scratch
        wgroups = (n_whole + 31) // 32
        has_w = n_whole > 0
        dmma = dmma + np.where(has_w, 6, 0)
        scalar0 = scalar0 + wgroups * 59 + np.where(has_w, 36, 0)
        scalar1 = scalar1 + wgroups * 66 + np.where(has_w, 36, 0)
        if two_level:
            # the coarse class: one column tile per k-step, half the W' table; per item with such sites the shift onto
            # the 128-point blocks (two DMMAs per non-zero 8 x 4 block of the upper-triangular S, one multiply each,
            # 4 KM/8 additions into the moment scratch + KM/4 for the two accumulator sets)
            cgroups = (n_coarse + 31) // 32
            shift = sum(1 for mt in range(KM // 8) for ks in range(KM // 4) if 4 * ks + 3 >= 8 * mt)
            has = (n_coarse > 0) | has_w
            dmma = dmma + (KM // 8) * ksteps(n_coarse, 32) + np.where(has, 2 * shift, 0)
            scalar0 = scalar0 + cgroups * (30 + KM + NB // 2) + np.where(has, shift + 6 * (KM // 8), 0)
            scalar1 = scalar1 + cgroups * (30 + KM + NB // 2 + 27) + np.where(has, shift + 6 * (KM // 8), 0)
    else:
        n_all = n_slow + n_fast
        groups = (n_all + 31) // 32
        dmma = 4 * RH * ksteps(n_all, 32)
        per_group = 45 + 32 + 8 * RH
        scalar0 = groups * per_group + 16 * RH
        scalar1 = groups * (per_group + 28) + 16 * RH
    dmma_total = int(dmma.sum()) * n_tiles
    scalar_total = int(scalar0.sum()) + int(scalar1.sum()) * (n_tiles - 1)
    return {"dmma": dmma_total, "scalar": scalar_total, "lane_ops": 256 * dmma_total + 32 * scalar_total,
            "n_active": int(n_slow.sum() + n_fast.sum() + sum(int(x.sum()) for x in (n_coarse, n_whole) if x is not None))}


def executed_f32_ops(offsets, rates, T):
    """FP32 mode (profile_f32_kernel): FFMA lane-operations of one launch.  Per slow site (rate < 0.009765625) 128 for
    the moment outer product + 22 for its moments and block weights (one lane); per faster site 64 on each of the 32
    lanes; per item 16 x 32 per lane for the moments -> times Horner evaluation + the reduction tree."""
    n_tiles = max(1, (int(T) + 1023) // 1024)
    off = np.asarray(offsets, dtype=np.int64)
    MOM_BOUND["f32"] = 0.009765625
    n_slow, n_fast = np.zeros(off.size - 1, dtype=np.int64), np.zeros(off.size - 1, dtype=np.int64)
    l0, chunk = 0, 1 << 26
    L = off.size - 1
    while l0 < L:
        l1 = min(L, max(int(np.searchsorted(off, off[l0] + chunk, side="right")) - 1, l0 + 1))
        r = np.asarray(rates[off[l0]:off[l1]])
        mag = np.abs(r)
        active = (mag > 1.5717277847026288e-162) & (mag <= 3.3519519824856493e+153)
        slow = active & (r > 0.0) & (r < 0.009765625)
        seg, lens = off[l0:l1] - off[l0], off[l0 + 1:l1 + 1] - off[l0:l1]
        full = np.nonzero(lens > 0)[0]
        if full.size:
            n_slow[l0 + full] = np.add.reduceat(slow, seg[full], dtype=np.int64)
            n_fast[l0 + full] = np.add.reduceat(active & ~slow, seg[full], dtype=np.int64)
        l0 = l1
    lane_ops = int(150 * n_slow.sum() + 64 * 32 * n_fast.sum() + (16 * 32 * 32 + 4 * 31 * 32) * (off.size - 1)) * n_tiles
    return {"dmma": 0, "scalar": 0, "lane_ops": lane_ops, "n_active": int(n_slow.sum() + n_fast.sum()), "fp32": True}


# --------------------------------------------------------------------------------------------
def workload_name(w):
    from tapir_b200 import synthetic
    c = synthetic.CONFIGS[w]
    if c["kind"] == "uniform":
        return "%s: synthetic %d loci x %d sites, %d-point grid, 4 epochs, 4 times" % (w, c["n_loci"], c["n_sites"], c["T"])
    return "%s: synthetic %d loci x ~%d sites (ragged %d-%d), %d-point grid, 4 epochs, 4 times" % (
        w, c["n_loci"], c["mean_sites"], c["smin"], c["smax"], c["T"])


def run_reference(args, rank, world):
    """--impl reference: the CPU path alone, rank 0 only."""
    if rank != 0:
        return
    from tapir_b200 import synthetic
    cfg = synthetic.CONFIGS[args.workload]
    # a bounded prefix of the batch is all this arm ever touches (block-seeded generation: the same loci as the full batch)
    if cfg["kind"] == "uniform" and cfg["n_loci"] * cfg["n_sites"] > synthetic.BLOCKED_ABOVE:
        off, rates, T = synthetic.build(args.workload, loci=(0, min(cfg["n_loci"], 16384)))
    else:
        off, rates, T = synthetic.build(args.workload)
    times, epochs = synthetic.TIMES, synthetic.EPOCHS
    pool, workers = make_pool(T, times, epochs)
    L = off.size - 1
    try:
        # size one step for ~8 s so that steps + warmup finish within a few minutes
        cpu_reference_pass(off, rates, T, min(L, workers), pool)           # fork + import warm-up
        n0 = min(L, 4 * workers)
        dt0, _, _ = cpu_reference_pass(off, rates, T, n0, pool)
        n = int(min(L, max(n0, n0 * 8.0 / max(dt0, 1e-3))))
        for _ in range(args.warmup):
            cpu_reference_pass(off, rates, T, min(n, workers), pool)
        tot_t = tot_e = 0.0
        for _ in range(args.steps):
            dt, evals, _ = cpu_reference_pass(off, rates, T, n, pool)
            tot_t += dt
            tot_e += evals
    finally:
        pool.close()
        pool.join()
    value = tot_e / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "sample_loci_per_step": n, "timing": "host wall clock"},
        "loci_per_s": n * args.steps / tot_t,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": cpu_kind(),
                         "sample": "%d of %d loci per step, Pool(%d), %s" % (n, cfg["n_loci"], workers, cpu_how())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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
        # bounded by content, not only by count: at most ~1M sites of rate files and ~1M sqlite rows
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
        names = ["locus-%d.rates" % l for l in range(n_db)]
        t_db = {}
        for with_net in (True, False):
            conn, c = db.create_probe_db(os.path.join(d, "pi_%d.sqlite" % with_net))
            t0 = time.perf_counter()
            db.insert_batch(conn, c, names, sub, times, epochs, with_net=with_net)
            conn.commit()
            t_db[with_net] = time.perf_counter() - t0
            conn.close()
        rows = n_db * (1 + T + len(times) + len(epochs))
        return {"ingest_files": n_files, "ingest_native_files_per_s": n_files / t_native, "ingest_native_matches": ok,
                "ingest_json_load_files_per_s": n_files / t_py, "sqlite_loci": n_db, "sqlite_rows_per_s": rows / t_db[True],
                "sqlite_loci_per_s": n_db / t_db[True], "sqlite_loci_per_s_without_net_table": n_db / t_db[False],
                "note": "host side of the drop-in, outside the timed region: tapir_ingest_rates (C++ threads) vs json.load "
                        "per file; bulk executemany writer with the reference's schema, with and without the "
                        "T-rows-per-locus net table (--no-net-table)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# --------------------------------------------------------------------------------------------
# parity gate: sampled loci of the timed batch against the CPU oracle (outside the timed region)
# --------------------------------------------------------------------------------------------
def parity_sample(L, n=32, seed=12345):
    rng = np.random.default_rng(seed)
    idx = np.unique(np.concatenate([[0, L - 1], rng.integers(0, L, size=max(0, n - 2))]))
    return idx.astype(np.int64)


def oracle_for(offsets, rates, idx, T, epochs):
    """CPU oracle (QUADPACK port) of the loci `idx`: (pi_net [n,T], integral [n,E], abserr [n,E])."""
    from oracle import capi
    from tapir_b200.engine import shard_csr
    off_s, rates_s = shard_csr(offsets, rates, idx)
    return capi.batch(off_s, rates_s, T, epochs)


def rel_err(a, b, floor=1e-290):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0


def parity_check(want, pi_net_rows, small_rows, tol=1e-9):
    """want = oracle tuple for the sampled loci; the device's rows for the same loci."""
    pi_net, si, se = want
    out = {"n": int(si.shape[0]),
           "max_rel_pi_net": rel_err(pi_net_rows, pi_net),
           "max_rel_integral": rel_err(small_rows["integral"], si),
           "max_rel_abserr": rel_err(small_rows["abserr"], se),
           "argmax_epoch_identical": bool(np.array_equal(small_rows["argmax_epoch"], np.argmax(si, axis=1))),
           "argmax_t_identical": bool(np.array_equal(small_rows["argmax_t"], np.argmax(pi_net, axis=1)))}
    out["ok"] = bool(max(out["max_rel_integral"], out["max_rel_pi_net"]) <= tol and out["argmax_epoch_identical"]
                     and (out["argmax_t_identical"] or tol > 1e-9))
    out["gate"] = ("sampled loci of the timed batch vs oracle/quadpack_port.c: pi_net, integral <= %g relative "
                   "(%s mode), identical argmax epoch%s" % (tol, "FP64" if tol <= 1e-9 else "FP32", " and argmax t" if tol <= 1e-9 else ""))
    return out


# --------------------------------------------------------------------------------------------
def timed_steps(eng, sb, steps, warmup, barrier):
    """Attribution pass (kernels serialised on one stream: exact per-kernel times) and the timed steps proper."""
    eng.set_serial_kernels(True)
    for _ in range(max(warmup, 3)):
        sb.step()
    k = {"profile": 0.0, "quad": 0.0, "finalize": 0.0, "serialised_step": 0.0}
    for _ in range(steps):
        eng.flush_l2()
        tm = sb.step()
        k["profile"] += tm.profile_ms
        k["quad"] += tm.quad_ms
        k["finalize"] += tm.finalize_ms
        k["serialised_step"] += tm.total_ms
    eng.set_serial_kernels(False)
    for _ in range(max(warmup, 3)):
        sb.step()
        sb.gather_step()
    barrier()
    dev_ms, gather_ms, launches = 0.0, 0.0, 0
    for _ in range(steps):
        eng.flush_l2()                      # untimed: the step starts with a cold L2
        tm = sb.step()                      # CUDA events on the library stream, synchronised on both sides
        dev_ms += tm.total_ms
        launches += tm.launches
        if sb.mode == "nccl":
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sb.gather_step()
            e1.record()
            torch.cuda.synchronize()
            gather_ms += e0.elapsed_time(e1)
    barrier()
    k = {n: v / steps for n, v in k.items()}
    k["gather"] = gather_ms / steps
    return (dev_ms + gather_ms) / steps, k, launches, tm


class SoloBatch:
    """A whole batch on this rank's GPU, no gather (N = 1, and the one-shard-per-rank weak-scaling side config)."""
    mode = "none"

    def __init__(self, eng, off, rates, T, times, epochs):
        from tapir_b200.engine import pinned_empty
        self.eng, self.T, self.times, self.epochs = eng, T, times, epochs
        self.offsets = pinned_empty(off.shape, np.int64)
        self.offsets[:] = off
        self.rates = pinned_empty(rates.shape, np.float64)
        self.rates[:] = rates
        eng.upload(self.offsets, self.rates, T, times, epochs, want_profile=True)
        self.loads = np.array([rates.size])
        self.parts = [np.arange(off.size - 1)]

    def step(self):
        return self.eng.compute()

    def gather_step(self):
        pass

    def run_host(self, out=None, pinned_results=True):
        return self.eng.run(self.offsets, self.rates, self.T, self.times, self.epochs, want_profile=True, out=out,
                            pinned_results=pinned_results)

    def close(self):
        pass


def side_config(eng, name, steps, e2e_steps, dist_state, gather):
    """A secondary workload, device-timed (and optionally end to end), reported under other_configs."""
    from tapir_b200 import sharded, synthetic
    rank, world, barrier, allmax = dist_state
    times, epochs = synthetic.TIMES, synthetic.EPOCHS
    if name == "cfg2_weak":
        cfg = dict(synthetic.CONFIGS["cfg2"])
        cfg["seed"] = cfg["seed"] * 1000 + rank
        synthetic.CONFIGS["_shard"] = cfg
        off, rates, T = synthetic.build("_shard")
        label = workload_name("cfg2") + " per GPU (weak scaling: one such shard per rank, as in round 1)"
        sb = SoloBatch(eng, off, rates, T, times, epochs)
        total_sites, total_loci = rates.size * world, (off.size - 1) * world
    else:
        off, rates, T = synthetic.build(name)
        label = workload_name(name)
        sb = sharded.ShardedBatch(eng, off, rates, T, times, epochs, gather=gather, want_profile=True)
        total_sites, total_loci = rates.size, off.size - 1
    step_ms, k, _, tm = timed_steps(eng, sb, steps, 3, barrier)
    step_ms = allmax(step_ms)
    rec = {"workload": label, "ms_per_step": step_ms, "value": total_sites * T / (step_ms * 1e-3), "unit": UNIT,
           "loci_per_s": total_loci / (step_ms * 1e-3), "steps": steps, "kernel_ms": k,
           "kernel": ("profile_mom_kernel<%d, %d moments>" % (tm.profile_tile_points // 256, tm.profile_moment_form)
                      if tm.profile_moment_form else "profile_kernel<%d>" % (tm.profile_tile_points // 256)),
           "epochs_from_moments": tm.moment_epochs}
    if world > 1 and name != "cfg2_weak":
        rec["partition"] = {"sites_per_rank_max_over_mean": float(sb.loads.max() / sb.loads.mean()),
                            "loci_per_rank": [int(len(p)) for p in sb.parts], "how": "partition_loci: longest-processing-time greedy"}
    if e2e_steps:
        out = sb.run_host()
        sb.run_host(out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sb.run_host(out=out)
        dt = allmax((time.perf_counter() - t0) / e2e_steps)
        rec["e2e"] = {"value": total_sites * T / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": e2e_steps}
        del out
    sb.close()
    return rec


def run_ours(args, rank, world, local_rank):
    from tapir_b200 import sharded, synthetic
    from tapir_b200.engine import Engine, bind_host_thread_near

    times, epochs = synthetic.TIMES, synthetic.EPOCHS
    t_build = time.perf_counter()
    off_np, rates_np, T = synthetic.build(args.workload)          # every rank builds the SAME full batch
    t_build = time.perf_counter() - t_build
    L, nnz = off_np.size - 1, rates_np.size

    # parity sample and its oracle, and the CPU arm, before any CUDA call (fork safety)
    sample = parity_sample(L)
    want = oracle_for(off_np, rates_np, sample, T, epochs) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(off_np, rates_np, T, times, epochs)

    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    numa_bound = bind_host_thread_near(local_rank) if world > 1 else False   # after the CPU pool (N = 1 keeps all cores)
    if args.fp32:
        os.environ["TAPIR_FP32_MODE"] = "1"               # read when the context is created: every batch in FP32 mode
    eng = Engine(local_rank)

    # ---- ONE batch, split over the ranks by the size-balanced locus partition ----
    t_shard = time.perf_counter()
    sb = sharded.ShardedBatch(eng, off_np, rates_np, T, times, epochs, gather=args.gather, want_profile=True)
    t_shard = time.perf_counter() - t_shard
    my_sites = int(sb.rates.size)
    hs_prefix = None
    if rank == 0 and world == 1 and not args.no_host_side:
        n_hs = min(L, 512)
        hs_prefix = (np.array(off_np[:n_hs + 1]), np.array(rates_np[:off_np[n_hs]]))
    del rates_np                                               # the partition lives in pinned memory now

    sampler = ClockSampler(local_rank, period_s=0.001)
    if rank == 0:
        sampler.start()
    step_ms, kernel_ms, launches, tm = timed_steps(eng, sb, args.steps, args.warmup, barrier)
    per_rank_ms = [step_ms]
    if dist is not None:
        per_rank_ms = [None] * world
        dist.all_gather_object(per_rank_ms, step_ms)
    step_ms = allmax(step_ms)

    # ---- correctness of exactly what was timed: the resident batch's results ----
    out = eng.download(pinned_results=True)                    # this rank's rows (profile included), pinned: reused below
    gathered = sb.collect(local=out)                           # rank 0: every locus, original order (barrier inside)
    where = np.empty(L, dtype=np.int64)
    for r, idx in enumerate(sb.parts):
        where[idx] = r
    rows = {int(l): np.array(out.pi_net[int(np.searchsorted(sb.mine, l))]) for l in sample if where[l] == rank}
    my_sum = float(np.sum(out.integral))
    if dist is not None:
        allrows, sums = [None] * world, [None] * world
        dist.all_gather_object(allrows, rows)
        dist.all_gather_object(sums, my_sum)
    else:
        allrows, sums = [rows], [my_sum]
    parity = gather_check = checksum = None
    if rank == 0:
        merged = {}
        for d in allrows:
            merged.update(d)
        pi_rows = np.stack([merged[int(l)] for l in sample])
        parity = parity_check(want, pi_rows, {n: gathered[n][sample] for n in ("integral", "abserr", "argmax_epoch", "argmax_t")},
                              tol=1e-5 if args.fp32 else 1e-9)
        checksum = float(np.sum(gathered["integral"]))         # original locus order: the same number at every N
        if world > 1:
            ok = np.isclose(checksum, float(np.sum(sums)), rtol=1e-12)
            gather_check = "ok" if ok else "MISMATCH %r vs %r" % (checksum, float(np.sum(sums)))
    del gathered

    # ---- end to end through the public call: pinned host in, pinned host out, per-locus vectors at rank 0 ----
    for _ in range(2):
        sb.run_host(out=out)
        if world > 1:
            sb.collect()
    h0, d0 = eng.transfer_bytes()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sb.run_host(out=out)
        if world > 1:
            sb.collect()                                       # barrier + rank 0 reads the gather buffer, un-permutes
    e2e_s = allmax((time.perf_counter() - t0) / args.steps)
    h1, d1 = eng.transfer_bytes()
    clocks = sampler.stop() if rank == 0 else None
    e2e_checksum = float(np.sum(out.integral))
    # PCIe context for the e2e number: the D2H copy of this rank's results on its own (pinned destination)
    eng.upload(sb.offsets, sb.rates, T, times, epochs, want_profile=True)
    eng.compute()
    eng.download(out=out)
    _, dd0 = eng.transfer_bytes()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.download(out=out)
    d2h_s = (time.perf_counter() - t0) / 3
    _, dd1 = eng.transfer_bytes()
    d2h_gbs = (dd1 - dd0) / 3 / d2h_s / 1e9
    peak_tf, _ = eng.fp64_peak()
    hs_out = None
    if hs_prefix is not None:
        hs_out = eng.run(hs_prefix[0], hs_prefix[1], T, times, epochs, want_profile=True)
    del out

    # executed FP64 work of rank 0's profile launch (counted from its own partition)
    ex = None
    if rank == 0 and tm.profile_fp32:
        ex = executed_f32_ops(sb.offsets, sb.rates, T)
    elif rank == 0:
        ex = executed_fp64_ops(sb.offsets, sb.rates, T, tm.profile_tile_points, tm.profile_moment_form, tm.moment_epochs)
    mode, loads, parts_len = sb.mode, sb.loads, [int(len(p)) for p in sb.parts]
    sb.close()
    del sb

    # ---- secondary workloads (device-timed, few steps) ----
    others = {}
    if not args.no_other_configs:
        dist_state = (rank, world, barrier, allmax)
        if world == 1:
            others["cfg2"] = side_config(eng, "cfg2", max(3, min(args.steps, 10)), 5, dist_state, args.gather)
            others["cfg3"] = side_config(eng, "cfg3", 3, 0, dist_state, args.gather)
        else:
            others["cfg3"] = side_config(eng, "cfg3", 3, 0, dist_state, args.gather)
            others["cfg2_weak"] = side_config(eng, "cfg2_weak", max(3, min(args.steps, 10)), 0, dist_state, args.gather)

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    total_evals = nnz * T
    peaks, peak_src = measured_peaks()
    prof_s = kernel_ms["profile"] * 1e-3
    kernel_name = ("profile_mom_kernel<%d, %d moments>" % (tm.profile_tile_points // 256, tm.profile_moment_form)
                   if tm.profile_moment_form else "profile_kernel<%d>" % (tm.profile_tile_points // 256))
    if tm.profile_fp32:
        kernel_name = "profile_f32_kernel (1,024-point tiles, 16 moments, FP32)"
    direct_tf = 2.0 * ex["n_active"] * T / prof_s / 1e12        # one FMA per (active site, timepoint): the direct formulation
    exec_tf = 2.0 * ex["lane_ops"] / prof_s / 1e12             # what the kernel issues to the FP64 (FP32 mode: FP32) pipe
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12                 # nominal FFMA rate: 128 lanes per SM and clock
    line = {
        "metric": METRIC, "value": total_evals / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if tm.profile_fp32 else "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "loci": L, "loci_per_gpu": parts_len,
                   "grid": T, "epochs": epochs, "times": times,
                   "cache": "L2 flushed (256 MB memset) before every timed step",
                   "timing": "CUDA events on the launching stream per step, max over ranks",
                   "parallelism": ("ONE batch split over %d GPUs by partition_loci (size-balanced LPT) + shard_csr; per-locus "
                                   "vectors gathered to rank 0 by %s and un-permuted to the original locus order there" % (
                                       world, "NVLink peer-memory push fused after finalize (CUDA IPC)" if mode == "push"
                                       else "one NCCL gather per step")) if world > 1 else "single GPU",
                   "partition": {"sites_per_rank_max_over_mean": float(loads.max() / loads.mean())}},
        "loci_per_s": L / (step_ms * 1e-3), "ms_per_step_per_rank": [round(float(x), 4) for x in per_rank_ms],
        "kernel_ms": dict(kernel_ms, note="per-kernel times from a separate pass with the kernels serialised on one stream "
                                          "(rank 0); ms_per_step is the shipped two-stream configuration"),
        "roofline": {
            "bound": "tensor", "pipe": "FP64 (DMMA sub-pipe of the tensor pipe; shared with DFMA on B200)",
            "kernel": kernel_name, "epochs_from_moments": tm.moment_epochs,
            # executed: DMMA.8x8x4 (256 FMA lanes) and scalar FP64 warp instructions (32 lanes) the launch issues, counted
            # from the batch by the kernel's own loop structure (checked against ncu's opcode counts in the tests)
            "achieved": exec_tf, "peak": fp32_peak if tm.profile_fp32 else peak_tf, "unit": "TFLOP/s",
            "frac": exec_tf / (fp32_peak if tm.profile_fp32 else peak_tf) if peak_tf else None,
            "fp32_mode": bool(tm.profile_fp32),
            "executed": {"dmma_instructions": ex["dmma"], "scalar_fp64_warp_instructions": ex["scalar"],
                         "fma_lane_ops": ex["lane_ops"], "launch": "rank 0's partition, %d sites" % my_sites},
            "direct_form_equivalent": {
                "achieved": direct_tf, "frac": direct_tf / peak_tf if peak_tf else None,
                "note": "2 flop per (site with a non-zero finite rate, grid point) -- the work of the reference's "
                        "formulation -- over the same kernel time; the moment form executes about a quarter of it, so "
                        "this ratio is an algebra speed-up and may exceed 1, not a pipe utilisation"},
            "traffic": profile_traffic(args.workload),
            # the committed ncu capture of the same kernel: the FP64 pipe is not what bounds the moment form (the L1
            # data pipe -- the shared-memory transposition of the group tables -- is busier); see DESIGN.md 3.1a
            "ncu_utilisation": profile_ncu_utilisation(args.workload),
            "peak_source": "FP64 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has HBM and bf16 entries only; "
                           "tools/microbench/fp64_pipes.cu: DMMA 37.0, DFMA 36.6 TFLOP/s, one shared pipe)",
            "active_site_fraction": ex["n_active"] / max(my_sites, 1),
            "quad": {"ms": kernel_ms["quad"], "note": "quad_scan_kernel + quad_kernel + quad_fix_kernel: rate statistics and "
                                                      "the sites the moment functional does not serve"},
            # the same evals/s against the ceiling of the reference's formulation, one FP64 exp per
            # (site, timepoint): ~17 pipe operations for a branch-free exp + 3 for a*t*e, on the same measured pipe
            "naive_one_exp_per_eval": {"ops_per_eval": 20, "ceiling_evals_per_s": peak_tf * 1e12 / 2.0 / 20.0 * world,
                                       "value_over_ceiling": (total_evals / (step_ms * 1e-3)) / (peak_tf * 1e12 / 2.0 / 20.0 * world)
                                       if peak_tf else None},
            "hbm": {"achieved_gbs": 8.0 * my_sites * max(1, -(-T // tm.profile_tile_points)) / prof_s / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src,
                    "note": "rate stream of the profile kernel, 8 B per site per time tile"},
        },
        "e2e": {"value": total_evals / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": (h1 - h0) // args.steps, "d2h_bytes_per_step": (d1 - d0) // args.steps,
                "bytes_are": "rank 0's (its partition's rates + work lists in, its profile rows + per-locus vectors out%s)" % (
                    "; plus the gathered vectors of all ranks" if world > 1 else ""),
                "loci_per_s": L / e2e_s, "timing": "host wall clock around tapir_cuda_run, pinned buffers, max over ranks",
                "d2h_alone_ms": d2h_s * 1e3, "d2h_gbs": d2h_gbs, "host_thread_bound_to_gpu_numa_node": numa_bound,
                "note": "chunks of loci (13 MB of traffic each, a quarter- and a half-size one first) rotate over up to eight "
                        "lanes (H2D / kernels / D2H of consecutive chunks overlap); bounded below by the D2H copy of the profile"},
        "parity": parity,
        "gpu_launches": launches, "gather": {"mode": mode, "check": gather_check},
        "clocks": clocks, "checksum": checksum, "e2e_checksum_rank0": e2e_checksum,
        "host_prep": {"build_batch_s": t_build, "partition_shard_upload_s": t_shard,
                      "note": "untimed: synthetic batch generation; partition_loci + shard_csr into pinned memory + first upload"},
        "other_configs": others,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if hs_out is not None:
        line["host_side"] = host_side_rates(hs_prefix[1], hs_prefix[0], T, times, epochs, hs_out)
    if parity is not None and not parity["ok"]:
        sys.stderr.write("PARITY GATE FAILED: %r\n" % (parity,))
    print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-host-side", action="store_true", help="skip the (untimed) ingest / sqlite sample")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the secondary workloads")
    ap.add_argument("--gather", default="push", choices=["push", "nccl"], help="N > 1: how per-locus vectors reach rank 0")
    ap.add_argument("--fp32", action="store_true", help="TAPIR_FP32_MODE: profile and slow sites' epoch sums in single "
                                                        "precision (parity gate 1e-5); the default and the headline are FP64")
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
