set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_v15.err | grep '^{' > gpurun_out/bench_v15.json; echo rc=$?
python bench.py --genome-bp 3100000000 --reads 100000 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c3b.err | grep '^{' > gpurun_out/bench_c3b.json; echo rc=$?
python - <<PY
import json
for f in ("bench_v15","bench_c3b"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"])
PY
