#!/bin/bash
tag=${1:-r02r}
mkdir -p gpurun_out
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r=d["roofline"]
    print("  step %.4f ms value %.3e kernel_ms %s %s frac %.3f e2e %.3f parity %s" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, r["kernel"], r["frac"], d["e2e"]["ms_per_step"], d["parity"]["ok"]))
except Exception as e:
    print("  no result", e)
PY
}
for w in cfg5 cfg3; do
 for rh in 4 8; do
  f=gpurun_out/${tag}_bench_${w}_rh${rh}_k16.json
  TAPIR_CUDA_RH=$rh TAPIR_MOM_TERMS=16 timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err; echo "bench $w RH=$rh K=16 rc=$?"; summ $f
 done
done
f=gpurun_out/${tag}_bench_cfg2.json
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err; echo "bench cfg2 rc=$?"; summ $f
