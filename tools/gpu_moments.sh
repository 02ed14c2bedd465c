#!/bin/bash
# 16 against 32 moments per block on configs 2, 3, 5 (TAPIR_MOM_TERMS forced), and the model's own choice
tag=${1:-mom}
out=gpurun_out/${tag}_moments_sweep.txt
echo "# 16 vs 32 moments per 128-point block (TAPIR_MOM_TERMS), bench.py --workload W --no-other-configs, B200, final build of round 2; 'auto' = choose_moments" > $out
for w in cfg2 cfg5 cfg3; do
  for k in 32 16 auto; do
    f=gpurun_out/${tag}_mom_${w}_${k}.json
    steps=5; [ "$w" = "cfg2" ] && steps=20
    if [ "$k" = "auto" ]; then unset TAPIR_MOM_TERMS; else export TAPIR_MOM_TERMS=$k; fi
    timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err
    python - $f $w $k >> $out <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    k=d["kernel_ms"]
    print("%-5s K=%-4s step %8.4f ms  profile %8.4f  quad %8.4f  finalize %7.4f  %s  pi_net err %.1e integral err %.1e" % (sys.argv[2], sys.argv[3], d["ms_per_step"], k["profile"], k["quad"], k["finalize"], d["roofline"]["kernel"], d["parity"]["max_rel_pi_net"], d["parity"]["max_rel_integral"]))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "no result", e)
PY
  done
done
unset TAPIR_MOM_TERMS
cat $out
