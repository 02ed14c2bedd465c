#!/usr/bin/env python
"""Profile-kernel time over grid sizes and tile variants (direct form RH = 1, 2, 4, 8 / moment form RH = 4, 8), the data the
tile-size model in tapir_cuda.cu (tile_cost) is calibrated against.  Run on the GPU box:
    python tools/tile_sweep.py > gpurun_out/r02_tile_sweep.txt
Each line: T, sites per locus, variant, profile kernel ms (kernels serialised, L2 flushed, median of 5)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tapir_b200 import synthetic                      # noqa: E402
from tapir_b200.engine import Engine                  # noqa: E402

GRIDS = [200, 256, 300, 400, 512, 600, 800, 1000, 1025, 1100, 1500, 2048, 2049, 2500, 3000, 4096, 5000]
VARIANTS = [("d1", 1, 0), ("d2", 2, 0), ("d4", 4, 0), ("d8", 8, 0), ("m4", 4, 1), ("m8", 8, 1)]


def main():
    shapes = [(5000, 500), (20000, 100), (1500, 2000)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        shapes = shapes[:1]
    print("# T sites variant profile_ms step_ms  (auto = the model's choice: tile points / moment form)")
    for n_loci, n_sites in shapes:
        for T in GRIDS:
            off, rates = synthetic.csr_uniform(7, n_loci, n_sites, T)
            row = {}
            for name, rh, mom in VARIANTS + [("auto", 0, 1)]:
                if rh:
                    os.environ["TAPIR_CUDA_RH"] = str(rh)
                else:
                    os.environ.pop("TAPIR_CUDA_RH", None)
                os.environ["TAPIR_PROF_MOM"] = str(mom)
                eng = Engine(0)
                try:
                    eng.upload(off, rates, T, synthetic.TIMES, synthetic.EPOCHS, want_profile=True)
                    eng.set_serial_kernels(True)
                    for _ in range(3):
                        eng.compute()
                    ms, tot = [], []
                    for _ in range(5):
                        eng.flush_l2()
                        tm = eng.compute()
                        ms.append(tm.profile_ms)
                        tot.append(tm.total_ms)
                    row[name] = (float(np.median(ms)), float(np.median(tot)), tm.profile_tile_points, tm.profile_moment_form)
                finally:
                    eng.close()
            best = min((v[0], k) for k, v in row.items() if k != "auto")
            for name in [v[0] for v in VARIANTS] + ["auto"]:
                p, t, tp, mf = row[name]
                extra = " <- best" if name == best[1] else ""
                if name == "auto":
                    extra = "  auto = %d%s  (%.2fx best)" % (tp, "m" if mf else "d", p / best[0])
                print("%5d %5d %-4s %8.4f %8.4f%s" % (T, n_sites, name, p, t, extra))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
