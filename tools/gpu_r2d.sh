#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("  step %.4f ms value %.3e kernel_ms %s e2e %.4f ms  %s" % (d["ms_per_step"], d["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, d["e2e"]["ms_per_step"], d["roofline"]["kernel"]))
except Exception as e:
    print("  no result", e)
PY
}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${tag}_pytest.log
timeout 600 python tools/diag_nmulti.py > gpurun_out/${tag}_diag.txt 2>&1; echo "diag rc=$?"; cat gpurun_out/${tag}_diag.txt | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-host-side > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench cfg2 rc=$?"; summ gpurun_out/${tag}_bench_cfg2.json
for w in cfg5 cfg3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-host-side > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; echo "bench $w rc=$?"; summ gpurun_out/${tag}_bench_${w}.json
done
