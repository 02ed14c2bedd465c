#!/bin/bash
# round 2: parity tests, smoke, benches of configs 2/5/3, quick tile sweep, ncu launch list + one full capture
tag=${1:-r02b}
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
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-host-side > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench cfg2 rc=$?"; summ gpurun_out/${tag}_bench_cfg2.json
for w in cfg5 cfg3; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-host-side > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; echo "bench $w rc=$?"; summ gpurun_out/${tag}_bench_${w}.json
done
timeout 600 python tools/tile_sweep.py quick > gpurun_out/${tag}_tile_sweep.txt 2> gpurun_out/${tag}_tile_sweep.err; echo "sweep rc=$?"; grep auto gpurun_out/${tag}_tile_sweep.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-host-side"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_mom_kernel|quad_scan_kernel|quad_kernel' -s 3 -c 3 -o gpurun_out/${tag}_prof -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
