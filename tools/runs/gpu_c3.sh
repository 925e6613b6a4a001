set -x
python bench.py --index partitioned --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_part1.json 2> gpurun_out/bench_part1.err; echo rc=$?
tail -c 600 gpurun_out/bench_part1.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_part1.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d.get("exchange"))
PY
free -g | head -2
timeout 1200 python bench.py --genome-bp 3100000000 --reads 100000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo rc=$?
tail -c 1500 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
