set -x
for R in 2000000 4000000; do
python bench.py --reads $R --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r$R.err | grep '^{' > gpurun_out/bench_r$R.json; echo rc=$?
done
python - <<PY
import json
for f in ("bench_r2000000","bench_r4000000"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["stages_ms_per_step"], d["roofline"]["frac"])
PY
