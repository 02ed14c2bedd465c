#!/bin/bash
# Weak-scaling runs of bench.py on one box: usage gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh tag 1 2 4 8'
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in "$@"; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_scale_n$n.json") if l.startswith("{")][0])
    print("  value %.4g evals/s  step %.4f ms  e2e %.3f ms  gather %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("gather")))
except Exception as e:
    print("  no result:", e)
PY
  grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${tag}_scale_n$n.err | tail -3 | cut -c1-300
done
