mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r08g_pytest_mom.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r08g_pytest_mom.log
for m in 1 1; do
TAPIR_PROF_MOM=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-host-side > gpurun_out/r08g_bench_m$m.json 2> gpurun_out/r08g_bench_m$m.err; echo "mom=$m rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r08g_bench_m$m.json") if l.startswith("{")][-1])
r=d["roofline"]
print("step %.4f ms  value %.3e  profile %.4f frac %.3f  quad %.4f ms  e2e %.4f checksum %.12g" % (d["ms_per_step"], d["value"], d["kernel_ms"]["profile"], r["frac"], r["quad_kernel"]["ms"], d["e2e"]["ms_per_step"], d["checksum"]))
PY
done
python tools/prof_trace.py cfg2 > gpurun_out/r08g_trace.txt 2>&1; tail -8 gpurun_out/r08g_trace.txt
