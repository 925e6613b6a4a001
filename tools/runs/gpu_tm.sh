set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tm.err | grep '^{' > gpurun_out/bench_tm.json; echo rc=$?
HRM_PROBE_TM=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_rm.err | grep '^{' > gpurun_out/bench_rm.json; echo rc=$?
HRM_PROBE_TM=1 python bench.py --genome-bp 46000000 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tmc2.err | grep '^{' > gpurun_out/bench_tmc2.json; echo rc=$?
python - <<PY
import json
for f in ("bench_tm","bench_rm","bench_tmc2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["value"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d["roofline"]["launch_ms"], d.get("stages_ms_per_step"))
    except Exception as e:
        print(f, "failed", e)
PY
