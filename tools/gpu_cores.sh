#!/bin/bash
# co-residency experiment: cap the profile kernel's one-warp CTAs per SM so that the integral kernels (second stream) fit beside it
tag=${1:-cores}; shift
for c in "$@"; do
  f=gpurun_out/${tag}_cap${c}.json
  TAPIR_PROF_CTAS_PER_SM=$c timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err
  python - $f $c <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("cap %s step %.4f ms kernel_ms %s parity %s" % (sys.argv[2], d["ms_per_step"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k not in ("note","gather")}, d["parity"]["ok"]))
except Exception as e:
    print(sys.argv[2], "no result", e)
PY
done
