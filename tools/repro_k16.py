import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TAPIR_MOM_TERMS"] = "16"
os.environ["TAPIR_CUDA_RH"] = sys.argv[1] if len(sys.argv) > 1 else "4"
from tapir_b200 import synthetic
from tapir_b200.engine import Engine, pack_csr
rng = np.random.default_rng(1)
T = 1000
loci = [synthetic.uce_rates(rng, n, T) for n in (500, 500, 37, 300)]
off, rates = pack_csr(loci, pinned=False)
eng = Engine(0)
eng.set_serial_kernels(True)
eng.upload(off, rates, T, synthetic.TIMES, synthetic.EPOCHS)
tm = eng.compute()
print("ok", tm.profile_tile_points, tm.profile_moment_form, tm.profile_ms, tm.quad_ms)
res = eng.download()
print(res.integral[:2])
