set -x
python -m pytest tests -m gpu -x -q -k "fused or scale or small" 2>&1 | tail -4
for S in 4096 2048; do
HRM_COLLECT_SLOTS=$S python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_f$S.err | grep '^{' > gpurun_out/bench_f$S.json; echo rc=$?
done
python bench.py --genome-bp 46000000 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_fc2.err | grep '^{' > gpurun_out/bench_fc2.json; echo rc=$?
python - <<PY
import json
for f in ("bench_f4096","bench_f2048","bench_f8192","bench_fc2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["value"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("stages_ms_per_step"))
    except Exception as e:
        print(f, "failed", e)
PY
