#!/usr/bin/env python
"""quad_kernel time against the number of epochs (config-2 rates; kernels serialised so the per-kernel events are exact).
usage (GPU box): python tools/quad_epochs.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tapir_b200 import synthetic
from tapir_b200.engine import Engine, pinned_empty

off_np, rates_np, T = synthetic.build("cfg2")
eng = Engine(0)
eng.set_serial_kernels(True)
offsets = pinned_empty(off_np.shape, np.int64); offsets[:] = off_np
rates = pinned_empty(rates_np.shape, np.float64); rates[:] = rates_np
all_epochs = [[32, 42], [88, 98], [95, 105], [164, 174], [10, 20], [50, 60], [120, 130], [140, 150]]
mixed = [[32, 42], [88, 98], [90, 110], [150, 174]]   # two half-lengths in one batch of four
for name, ep in [("E=%d equal widths" % e, all_epochs[:e]) for e in range(1, 9)] + [("E=4 mixed widths", mixed)]:
    eng.upload(offsets, rates, T, synthetic.TIMES, ep, want_profile=True)
    ms = []
    for _ in range(6):
        eng.flush_l2()
        ms.append(eng.compute().quad_ms)
    print("%-20s quad %.4f ms  (%.4f per epoch)" % (name, min(ms[1:]), min(ms[1:]) / len(ep)))
