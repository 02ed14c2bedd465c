#!/bin/bash
# One gpurun call: parity tests, smoke, a short bench, then the ncu launch list + one full capture.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 300 $CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_mom_kernel|quad_kernel' -s 2 -c 2 -o gpurun_out/${tag}_prof -f $CMD > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
