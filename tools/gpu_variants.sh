#!/bin/bash
# bench the library builds under variants/*.so one after the other (measurement helper); NCU=1 adds light ncu metrics
tag=${1:-var}; shift
mkdir -p gpurun_out
cp tapir_b200/libtapir_cuda.so /tmp/lib_orig.so
M=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for v in variants/*.so; do
  name=$(basename $v .so)
  cp $v tapir_b200/libtapir_cuda.so
  for w in ${@:-cfg5}; do
    f=gpurun_out/${tag}_${name}_${w}.json
    steps=5; [ "$w" = "cfg2" ] && steps=20
    timeout 600 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > $f 2> ${f%.json}.err
    python - $f $name $w <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("%s %s step %.4f ms kernel_ms %s parity %s" % (sys.argv[2], sys.argv[3], d["ms_per_step"], {k: round(v, 4) for k, v in d["kernel_ms"].items() if k not in ("note","gather")}, d["parity"]["ok"]))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "no result", e)
PY
    if [ -n "$NCU" ]; then
    timeout 600 ncu --metrics $M --clock-control none -k regex:profile_mom_kernel -s 3 -c 1 --csv --log-file gpurun_out/${tag}_${name}_${w}_icc.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs > /dev/null 2>&1
    python - gpurun_out/${tag}_${name}_${w}_icc.csv <<'PY'
import csv, sys
out=[]
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 12 and r[0].isdigit(): out.append("%s=%s" % (r[-3].split('.')[0].replace('sm__','').replace('smsp__','')[:34], r[-1]))
print("   ", "  ".join(out))
PY
    fi
  done
done
cp /tmp/lib_orig.so tapir_b200/libtapir_cuda.so
