#!/usr/bin/env python
"""Relative error of profile_mom_kernel (through Engine.run) against the CPU oracle on rate mixes that stress the
moment form: rates at the slow/fast threshold, mixed fast sites, several tile sizes.  usage (GPU box): python tools/mom_error.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tapir_b200 import synthetic
from tapir_b200.engine import Engine, pack_csr
from oracle import capi

eng = Engine(0)
rng = np.random.default_rng(7)
def rel(a, b): return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-290)))
cases = []
for T in (1000, 1024, 2048, 4096, 10000):
    cases.append(("UCE-like gamma rates, T=%d" % T, [synthetic.uce_rates(rng, 500, T) for _ in range(12)], T))
cases.append(("rates just below the threshold (0.0190-0.01953), T=1000", [rng.uniform(0.0190, 0.01953, 400) for _ in range(8)], 1000))
cases.append(("rates straddling the threshold (0.015-0.025), T=1000", [rng.uniform(0.015, 0.025, 400) for _ in range(8)], 1000))
cases.append(("slow + 40 fast sites (0.05-3), T=2048", [np.concatenate([rng.gamma(0.5, 0.002, 300), rng.uniform(0.05, 3.0, 40)]) for _ in range(8)], 2048))
cases.append(("all fast (0.02-1), T=1000", [rng.uniform(0.02, 1.0, 200) for _ in range(8)], 1000))
for name, loci, T in cases:
    off, rates = pack_csr(loci)
    res = eng.run(off, rates, T, [1, T // 2, T - 1], synthetic.EPOCHS, want_profile=True)
    pi_net, si, se = capi.batch(off, rates, T, synthetic.EPOCHS)
    am = bool(np.array_equal(res.argmax_t, np.argmax(pi_net, axis=1)))
    print("%-62s pi_net max rel err %.2e   argmax identical: %s" % (name, rel(res.pi_net, pi_net), am))
