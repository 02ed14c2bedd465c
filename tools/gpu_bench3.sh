#!/bin/bash
tag=${1:-r03d}
mkdir -p gpurun_out
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r=d["roofline"]
    print("  step %.4f ms value %.3e kernel_ms %s %s frac %.3f parity %s %.1e" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, r["kernel"], r["frac"], d["parity"]["ok"], d["parity"]["max_rel_pi_net"]))
except Exception as e:
    print("  no result", e)
PY
}
for w in ${2:-cfg5 cfg3 cfg2}; do
  steps=5; [ "$w" = "cfg2" ] && steps=20
  f=gpurun_out/${tag}_bench_${w}.json
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err; echo "bench $w rc=$?"; summ $f; tail -2 ${f%.json}.err | cut -c1-300
done
