#!/bin/bash
tag=${1:-r03c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${tag}_pytest.log | cut -c1-220
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r=d["roofline"]
    print("  step %.4f ms value %.3e kernel_ms %s %s frac %.3f e2e %.3f parity %s %.1e" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, r["kernel"], r["frac"], d["e2e"]["ms_per_step"], d["parity"]["ok"], d["parity"]["max_rel_pi_net"]))
except Exception as e:
    print("  no result", e)
PY
}
for w in cfg2 cfg5 cfg3; do
  steps=5; [ "$w" = "cfg2" ] && steps=20
  f=gpurun_out/${tag}_bench_${w}.json
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err; echo "bench $w rc=$?"; summ $f; tail -2 ${f%.json}.err | cut -c1-300
done
f=gpurun_out/${tag}_bench_cfg5_k32.json
TAPIR_MOM_TERMS=32 timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err; echo "bench cfg5 K=32 rc=$?"; summ $f
timeout 500 python tools/prof_trace.py cfg5 > gpurun_out/${tag}_trace_cfg5.txt 2>&1; grep -v "^  *0 " gpurun_out/${tag}_trace_cfg5.txt | grep -A30 "moment kernel" 
