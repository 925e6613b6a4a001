set -x
python -m pytest tests -m gpu -x -q -k "fused or scale or small" 2>&1 | tail -3 > gpurun_out/pt.txt; cat gpurun_out/pt.txt
for S in 1024 512; do
HRM_COLLECT_WARP_SLOTS=$S python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_w$S.err | grep '^{' > gpurun_out/bench_w$S.json; echo rc=$?
done
python bench.py --genome-bp 46000000 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_wc2.err | grep '^{' > gpurun_out/bench_wc2.json; echo rc=$?
python - <<PY
import json
for f in ("bench_w1024","bench_w512","bench_wc2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["value"], d["e2e"]["value"], d["stages_ms_per_step"]["filter"], d["stages_ms_per_step"]["shd"])
    except Exception as e:
        print(f, "failed", e)
PY
