#!/usr/bin/env python
"""Per-warp timeline of one profile_kernel launch (measurement only; TAPIR_PROF_TRACE stamps, see kernels.cuh).
usage (GPU box): python tools/prof_trace.py [workload] > gpurun_out/xyz.txt"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = os.path.join(ROOT, "gpurun_out", "prof_trace.bin")
os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ["TAPIR_PROF_TRACE"] = path
from tapir_b200 import synthetic
from tapir_b200.engine import Engine, pinned_empty

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
off_np, rates_np, T = synthetic.build(wl)
eng = Engine(0)
eng.set_serial_kernels(True)
offsets = pinned_empty(off_np.shape, np.int64); offsets[:] = off_np
rates = pinned_empty(rates_np.shape, np.float64); rates[:] = rates_np
eng.upload(offsets, rates, T, synthetic.TIMES, synthetic.EPOCHS, want_profile=True)
for _ in range(4):
    eng.flush_l2()
    tm = eng.compute()
print("profile kernel (traced build, serial): %.4f ms" % tm.profile_ms)
tr = np.fromfile(path, dtype=np.int64).reshape(-1, 32)
tr = tr[tr[:, 1] != 0]
W = tr.shape[0]
smid, t_in, t_claim, t_first, t_exit, n_done = tr[:, 0], tr[:, 1], tr[:, 2], tr[:, 3], tr[:, 12], tr[:, 13]
is_mom = bool((tr[:, 20] > 0).any())       # profile_mom_kernel stamps the phases of the second item instead
print("kernel: %s; warps %d (one-warp CTAs), SMs %d, items done %d" % (
    "profile_mom_kernel" if is_mom else "profile_kernel", W, np.unique(smid).size, n_done.sum()))
sm_t0 = {s: t_in[smid == s].min() for s in np.unique(smid)}
base = np.array([sm_t0[s] for s in smid])
def q(x): return "min %.0f  p50 %.0f  p90 %.0f  max %.0f" % (x.min(), np.percentile(x, 50), np.percentile(x, 90), x.max())
print("cycles since the SM's first warp entered:")
print("  entry            ", q(t_in - base))
if not is_mom:
    print("  first claim done ", q(t_claim - base))
    print("  first group popped", q(t_first - base))
print("  exit             ", q(t_exit - base))
print("items per warp histogram:", np.bincount(n_done))
if not is_mom:
    ends = tr[:, 4:12]
    dur = []
    for w in range(W):
        for k in range(min(n_done[w], 8)):
            dur.append((k, ends[w, k] - (t_in[w] if k == 0 else ends[w, k - 1])))
    dur = np.array(dur)
    for k in range(dur[:, 0].max() + 1):
        d = dur[dur[:, 0] == k, 1]
        print("  item %d of a warp: n %5d  duration %s" % (k, d.size, q(d)))
# per SM (the trace does not record which sub-partition a one-warp CTA landed on)
items_sm = np.array([n_done[smid == s].sum() for s in np.unique(smid)])
sm_span = np.array([(t_exit[smid == s] - sm_t0[s]).max() for s in np.unique(smid)])
print("items per SM: min %d  p50 %d  max %d (%.2f per sub-partition on average)" % (
    items_sm.min(), np.percentile(items_sm, 50), items_sm.max(), items_sm.mean() / 4.0))
print("SM busy span (cycles):", q(sm_span))
# inside the second item of a warp: groups 2..5 (profile_kernel's stamps; steady state, warps sharing a sub-partition)
g = tr[:, 16:32].reshape(-1, 4, 4)
ok = (n_done >= 2) & (g[:, :, 2].min(axis=1) > 0) & (tr[:, 20] == 0)
if ok.any():
    gg = g[ok]
    pop_to_built = (gg[:, :, 1] - gg[:, :, 0]).ravel()
    built_to_done = (gg[:, :, 2] - gg[:, :, 1]).ravel()
    done_to_pop = (gg[:, 1:, 0] - gg[:, :-1, 2]).ravel()
    print("second item, groups 2..5 (cycles):")
    print("  group popped -> tables built (exp chain, power tables, __syncwarp)", q(pop_to_built))
    print("  tables built -> contracted (128 DMMA = 2048 pipe cycles)          ", q(built_to_done))
    print("  contracted -> next group popped (rate rounds, queue)              ", q(done_to_pop))
    print("  whole group                                                       ", q((gg[:, 1:, 0] - gg[:, :-1, 0]).ravel()))

# profile_mom_kernel: phases of the warp's second item
m = (n_done >= 2) & (tr[:, 20] > 0)
if m.any():
    ph = tr[m]
    names = ["moment pass (slow sites)", "moments -> times", "fast-site pass", "epilogue"]
    print("moment kernel, second item of a warp (cycles): groups of the moment pass p50 %d, n_fast p50 %d" % (
        np.percentile(ph[:, 22], 50), np.percentile(ph[:, 24], 50)))
    for k, nm in enumerate(names):
        print("  %-28s" % nm, q(ph[:, 17 + k] - ph[:, 16 + k]))
    print("  %-28s" % "whole item", q(ph[:, 20] - ph[:, 16]))
    if (ph[:, 25:29] > 0).any():
        print("  of which, summed over the item's groups / sub-tiles:")
        for k, nm in enumerate(["groups of the moment pass (tables + DMMAs)", "moments -> times proper", "fast sites",
                                "epilogue (stores, maximum)"]):
            print("    %-44s" % nm, q(ph[:, 25 + k]))
        rounds = (ph[:, 17] - ph[:, 16]) - ph[:, 25]
        print("    %-44s" % "moment pass outside its groups (rate rounds)", q(rounds))
