set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu.txt; cat gpurun_out/r1_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 2>gpurun_out/r1_bench_human.err | grep '^{' > gpurun_out/r1_bench_human.json; echo rc=$?
python bench.py --genome-bp 46000000 --reads 1000000 --steps 5 --warmup 3 2>gpurun_out/r1_bench_chr21.err | grep '^{' > gpurun_out/r1_bench_chr21.json; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/r1_bench_reference.err | grep '^{' > gpurun_out/r1_bench_reference.json; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<PY
import json
for f in ("r1_bench_human","r1_bench_chr21","r1_bench_reference"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("stages_ms_per_step"), (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches"))
PY
