# one ncu --set full capture of the kernels matching $2 (regex), tag $1
tag=${1:-n}; pat=${2:-profile_kernel|quad_kernel}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 2 -c 2 -o gpurun_out/${tag}_prof -f $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
