#!/bin/bash
tag=${1:-r03k}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-250
bash tools/gpu_bench3.sh $tag
CMD5="python bench.py --workload cfg5 --steps 1 --warmup 3 --no-cpu-baseline --no-host-side --no-other-configs"
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
timeout 600 ncu --metrics $M --clock-control none -k regex:profile_mom_kernel -s 3 -c 1 --csv --log-file gpurun_out/${tag}_icc.csv $CMD5 > gpurun_out/${tag}_icc.log 2>&1; echo "ncu rc=$?"
python - gpurun_out/${tag}_icc.csv <<'PY'
import csv, sys
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 12 and r[0].isdigit(): print("  %-90s %s" % (r[-3], r[-1]))
PY
