#!/bin/bash
# round 2, first call: parity tests, smoke, A/B of the moment-functional integrals, tile sweep, large configs
tag=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${tag}_smoke.log
for mi in 1 0; do
  TAPIR_MOM_INT=$mi timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-host-side > gpurun_out/${tag}_bench_cfg2_momint$mi.json 2> gpurun_out/${tag}_bench_cfg2_momint$mi.err; echo "bench cfg2 momint=$mi rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_cfg2_momint$mi.json") if l.startswith("{")][-1])
    print("  step %.4f ms value %.3e kernel_ms %s e2e %.4f ms" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, d["e2e"]["ms_per_step"]))
except Exception as e:
    print("  no result", e)
PY
done
timeout 900 python tools/tile_sweep.py > gpurun_out/${tag}_tile_sweep.txt 2> gpurun_out/${tag}_tile_sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/${tag}_tile_sweep.err
for w in cfg5 cfg3; do
 for mi in 1 0; do
  TAPIR_MOM_INT=$mi timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-host-side > gpurun_out/${tag}_bench_${w}_momint$mi.json 2> gpurun_out/${tag}_bench_${w}_momint$mi.err; echo "bench $w momint=$mi rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_${w}_momint$mi.json") if l.startswith("{")][-1])
    print("  step %.4f ms value %.3e kernel_ms %s e2e %.4f ms" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, d["e2e"]["ms_per_step"]))
except Exception as e:
    print("  no result", e)
PY
 done
done
