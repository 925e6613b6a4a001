set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_r1.err | grep '^{' > gpurun_out/bench_r1.json; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r1.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["launch_ms"], d["stages_ms_per_step"], d.get("collect_ids_skipped_fraction"), d["cpu_baseline"]["value"], d["gpu_launches"])
PY
