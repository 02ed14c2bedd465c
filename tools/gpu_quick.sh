# parity tests + two short bench runs (tag = $1)
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
r=d["roofline"]
print("step %.4f ms  value %.3e  profile %.4f ms frac %.3f  quad %.4f ms  e2e %.4f ms" % (d["ms_per_step"], d["value"], 2*1750210000/ (r["achieved"]*1e12)*1e3, r["frac"], r["quad_kernel"]["ms"], d["e2e"]["ms_per_step"]))
PY
done
