#!/bin/bash
# Large BASELINE configs on one B200 (few steps; no CPU baseline): usage gpurun --timeout 1500 -- 'bash tools/gpu_large.sh tag cfg3 cfg5'
tag=$1; shift
mkdir -p gpurun_out
free -g | head -2
for w in "$@"; do
  echo "== $w"
  SECONDS=0; timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err
  echo "rc=$? wall=${SECONDS}s"; cut -c1-1800 gpurun_out/${tag}_bench_${w}.json; tail -3 gpurun_out/${tag}_bench_${w}.err | cut -c1-300
done
