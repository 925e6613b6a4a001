set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_h1.err | grep '^{' > gpurun_out/bench_h1.json; echo rc=$?
tail -c 800 gpurun_out/bench_h1.err
python bench.py --genome-bp 46000000 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c2.err | grep '^{' > gpurun_out/bench_c2.json; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_href.err | grep '^{' > gpurun_out/bench_href.json; echo rc=$?
python - <<PY
import json
for f in ("bench_h1","bench_c2","bench_href"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["value"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("stages_ms_per_step"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
