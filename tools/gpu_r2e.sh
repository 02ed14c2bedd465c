#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${tag}_pytest.log
SECONDS=0
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench (default) rc=$? wall=${SECONDS}s"; tail -5 gpurun_out/${tag}_bench.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
    print("  step %.4f ms value %.3e e2e %.3f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
    print("  kernel_ms", {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"})
    r=d["roofline"]; print("  roofline", r["kernel"], "frac %.3f direct %.3f" % (r["frac"], r["direct_form_equivalent"]["frac"]), r["executed"])
    print("  parity", d["parity"])
    print("  cpu", d.get("cpu_baseline"))
    print("  others", {k: (round(v["ms_per_step"], 4), "%.3e" % v["value"], v.get("e2e", {}).get("ms_per_step")) for k, v in d["other_configs"].items()})
    print("  host_prep", d["host_prep"], "host_side", {k: v for k, v in d.get("host_side", {}).items() if k != "note"})
except Exception as e:
    print("  no result", e)
PY
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "reference arm rc=$? wall=${SECONDS}s"; cut -c1-900 gpurun_out/${tag}_ref.json
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_mom_kernel|quad_scan_kernel|quad_kernel|finalize' -s 4 -c 4 -o gpurun_out/${tag}_prof -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${tag}_plain2.log | cut -c1-300
