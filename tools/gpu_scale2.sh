#!/bin/bash
# strong-scaling runs of bench.py on one box: usage gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale2.sh tag 2 4 8'
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi -L | head -8; nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1; nproc; free -g | head -2
for n in "$@"; do
  SECONDS=0
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
  fi
  echo "N=$n rc=$? wall=${SECONDS}s"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_scale_n$n.json") if l.startswith("{")][0])
    print("  value %.4g evals/s  step %.4f ms  e2e %.3f ms (%.3g evals/s) gather %s parity_ok %s checksum %r" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("gather"), d["parity"]["ok"], d["checksum"]))
    print("  kernel_ms", {k: round(v, 4) for k, v in d["kernel_ms"].items() if k != "note"}, "d2h_gbs %.1f" % d["e2e"]["d2h_gbs"])
    for k, v in d["other_configs"].items():
        print("  other", k, "step %.4f ms value %.4g" % (v["ms_per_step"], v["value"]), v.get("partition", ""))
except Exception as e:
    print("  no result:", e)
PY
  grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${tag}_scale_n$n.err | tail -4 | cut -c1-300
done
