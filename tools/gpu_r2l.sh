#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
CMD="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs --fp32"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_f32_kernel' -s 3 -c 1 -o gpurun_out/${tag}_prof_f32 -f $CMD > gpurun_out/${tag}_ncu_f32.log 2>&1
echo "ncu f32 rc=$?"
