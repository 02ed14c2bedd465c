"""Diagnostic (GPU): per (site, epoch) multi-panel decisions of the device path vs the QUADPACK port, every site as its own locus."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import capi
from tapir_b200.engine import Engine

T = 1000
rng = np.random.default_rng(T)
epochs = [[32, 42], [88, 98], [164, 174], [0, 64], [120, 136], [1, 2], [0, 200], [-3, 5], [60, 68]]
rng.gamma(0.5, 4.0 / T, 500)
lam = np.concatenate([rng.gamma(0.5, 0.004, 3000), rng.uniform(0.02, 2.0, 3000), rng.uniform(0.005, 0.05, 3000)])
off = np.arange(lam.size + 1, dtype=np.int64)
for mi in ("1", "0"):
    os.environ["TAPIR_MOM_INT"] = mi
    eng = Engine(0)
    res = eng.run(off, lam, T, [1], epochs, want_profile=False)
    eng.close()
    tot = 0
    for e, (a, b) in enumerate(epochs):
        integ, err, last = capi.epoch_sites(lam, a, b)
        # with one site per locus and one epoch column we cannot split n_multi per epoch; run epoch by epoch instead
    for e, (a, b) in enumerate(epochs):
        eng = Engine(0)
        r1 = eng.run(off, lam, T, [1], [[a, b]], want_profile=False)
        eng.close()
        integ, err, last = capi.epoch_sites(lam, a, b)
        want = (last > 1).astype(np.int64)
        bad = np.nonzero(r1.n_multi_panel != want)[0]
        relint = np.abs(r1.integral[:, 0] - integ) / (np.abs(integ) + 1e-290)
        relerr = np.abs(r1.abserr[:, 0] - err) / (np.abs(err) + 1e-290)
        print("MOM_INT=%s epoch %s: multi oracle %d device %d mismatches %d  max rel integral %.2e abserr %.2e" % (
            mi, (a, b), want.sum(), r1.n_multi_panel.sum(), bad.size, relint.max(), relerr.max()))
        for i in bad[:5]:
            print("    site rate %.17g: oracle last %d (res %.17g err %.17g) device multi %d res %.17g err %.17g" % (
                lam[i], last[i], integ[i], err[i], r1.n_multi_panel[i], r1.integral[i, 0], r1.abserr[i, 0]))
            print("       qk21:", capi.qk21(a, b, lam[i]), "errbnd", max(1.49e-8, 1.49e-8 * abs(integ[i])))
