set -x
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_validate.json 2> gpurun_out/bench_validate.err; echo rc=$?
tail -c 300 gpurun_out/bench_validate.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_validate.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "oneshot", d["e2e"]["one_shot_value"])
print("text", d["e2e_text"]["value"], "fastq", d["e2e_text"]["from_fastq_text"]["value"])
print("stages", d["stages_ms_per_step"])
print("roofline", d["roofline"]["frac"], d["roofline"]["frac_dram_side"], d["roofline"]["launch_ms"])
print({k:v for k,v in d.items() if k.startswith("parity")})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
PY
python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 60 > gpurun_out/bench_validate_ref.json 2> gpurun_out/bench_validate_ref.err; echo rc=$?
cut -c1-1500 gpurun_out/bench_validate_ref.json
