#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${tag}_pytest.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${tag}_smoke.log
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r=d["roofline"]
    print("  step %.4f ms value %.3e kernel_ms %s %s frac %.3f parity %s %.1e %.1e" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, r["kernel"], r["frac"], d["parity"]["ok"], d["parity"]["max_rel_pi_net"], d["parity"]["max_rel_integral"]))
except Exception as e:
    print("  no result", e)
PY
}
for w in cfg2 cfg5 cfg3; do
 for mode in "" "--fp32"; do
  steps=5; [ "$w" = "cfg2" ] && steps=20
  f=gpurun_out/${tag}_bench_${w}${mode/--/_}.json
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs $mode > $f 2> ${f%.json}.err; echo "bench $w $mode rc=$?"; summ $f; tail -2 ${f%.json}.err | cut -c1-300
 done
done
TAPIR_MOM_TERMS=16 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > gpurun_out/${tag}_bench_cfg2_k16.json 2> gpurun_out/${tag}_bench_cfg2_k16.err; echo "bench cfg2 K=16 rc=$?"; summ gpurun_out/${tag}_bench_cfg2_k16.json
