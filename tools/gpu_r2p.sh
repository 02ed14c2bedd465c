#!/bin/bash
tag=${1:-r02p}
mkdir -p gpurun_out
export TAPIR_MOM_TERMS=16
CMD="python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'quad_kernel|quad_scan_kernel|quad_fix' -s 9 -c 3 -o gpurun_out/${tag}_prof -f $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_ncu.log | cut -c1-200
