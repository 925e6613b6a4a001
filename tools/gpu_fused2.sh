set -x
for S in 2048 4096; do
HRM_COLLECT_SLOTS=$S python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_g$S.err | grep '^{' > gpurun_out/bench_g$S.json; echo rc=$?
done
python - <<PY
import json
for f in ("bench_g2048","bench_g4096"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["e2e"]["value"], d["stages_ms_per_step"]["filter"])
PY
