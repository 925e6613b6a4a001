# round 2: collection kernel A/B (bloom vs counting table) + parity of all collection paths
set -x
timeout 1500 python -m pytest tests/test_gpu_mapper.py -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_bloom.json 2> gpurun_out/bench_r2a_bloom.err; echo rc=$?
tail -c 300 gpurun_out/bench_r2a_bloom.err
HRM_COLLECT_RANGES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2a_ranges.json 2> gpurun_out/bench_r2a_ranges.err; echo rc=$?
python - <<PY
import json
for t in ("bloom","ranges"):
    d=json.load(open("gpurun_out/bench_r2a_%s.json"%t))
    print(t, d["value"], d["e2e"]["value"], d["stages_ms_per_step"], d["mapped_fraction"], d["candidates_per_read"])
PY
