set -x
python bench.py --reads 1000000 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_x.err | grep '^{' > gpurun_out/bench_x.json; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_x.json"))
print(d["value"], d["e2e"]["value"], d["stages_ms_per_step"]["filter"], d["stages_ms_per_step"]["shd"])
PY
