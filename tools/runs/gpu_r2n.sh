set -x
python bench.py --steps 3 --warmup 3 --check 0 --no-cpu-baseline > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo rc=$?
tail -c 500 gpurun_out/bench_r2n.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2n.json") if l.startswith("{")][0])
print("value", d["value"], "e2e", d["e2e"]["value"], "text", d["e2e_text"])
print("roofline", {k:v for k,v in d["roofline"].items() if k!="note"})
PY
python tools/run_100m.py --out gpurun_out/job_100m.json > gpurun_out/job_100m.log 2>&1; echo rc=$?
tail -3 gpurun_out/job_100m.log | cut -c1-1200
