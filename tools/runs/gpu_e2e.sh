set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pt.txt; cat gpurun_out/pt.txt
for C in 262144 131072 1000000; do
HRM_E2E_CHUNK=$C python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_e$C.err | grep '^{' > gpurun_out/bench_e$C.json; echo rc=$?
done
python - <<PY
import json
for f in ("bench_e262144","bench_e131072","bench_e1000000"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["e2e"]["value"])
PY
