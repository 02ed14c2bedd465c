#!/bin/bash
tag=${1:-r02i}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${tag}_smoke.log
for w in cfg2 cfg5 cfg3; do
 for k in 32 16 auto; do
  if [ "$k" = "auto" ]; then unset TAPIR_MOM_TERMS; else export TAPIR_MOM_TERMS=$k; fi
  steps=5; [ "$w" = "cfg2" ] && steps=20
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > gpurun_out/${tag}_bench_${w}_k$k.json 2> gpurun_out/${tag}_bench_${w}_k$k.err; echo "bench $w K=$k rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench_${w}_k$k.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print("  step %.4f ms value %.3e kernel_ms %s %s frac %.3f parity %s %.1e %.1e" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, r["kernel"], r["frac"], d["parity"]["ok"], d["parity"]["max_rel_pi_net"], d["parity"]["max_rel_integral"]))
except Exception as e:
    print("  no result", e)
PY
 done
done
unset TAPIR_MOM_TERMS
