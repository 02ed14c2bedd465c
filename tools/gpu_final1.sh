#!/bin/bash
# final single-GPU evidence of the round: tests, smoke, the driver's bench command, reference arm, ncu of config 2 and of the headline config's kernels
tag=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${tag}_smoke.log
SECONDS=0
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/${tag}_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
    print("  step %.4f ms value %.3e e2e %.3f ms (%.3e)" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"]))
    print("  kernel_ms", {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"})
    r=d["roofline"]; print("  roofline", r["kernel"], "frac %.3f direct %.3f" % (r["frac"], r["direct_form_equivalent"]["frac"]))
    print("  parity", {k: v for k, v in d["parity"].items() if k != "gate"})
    print("  cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"))
    print("  others", {k: (round(v["ms_per_step"], 4), "%.3e" % v["value"], v["kernel"], v.get("e2e", {}).get("ms_per_step")) for k, v in d["other_configs"].items()})
except Exception as e:
    print("  no result", e)
PY
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "reference arm rc=$? wall=${SECONDS}s"; cut -c1-400 gpurun_out/${tag}_ref.json
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs"
CMDL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs"
timeout 300 $CMDL > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMDL > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list (default workload) rc=$?"
timeout 300 $CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_mom_kernel|quad_scan_kernel|quad_kernel|finalize' -s 4 -c 4 -o gpurun_out/${tag}_prof_cfg2 -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full cfg2 rc=$?"
CMD5="python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs"
timeout 300 $CMD5 > gpurun_out/${tag}_plain5.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_mom_kernel|quad_scan_kernel|quad_kernel|finalize' -s 12 -c 4 -o gpurun_out/${tag}_prof_cfg5 -f $CMD5 > gpurun_out/${tag}_ncu_full5.log 2>&1
echo "ncu full cfg5 rc=$?"
