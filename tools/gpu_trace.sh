set -x
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_trace.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "text", d["e2e_text"]["value"], "fastq", d["e2e_text"]["from_fastq_text"]["value"])
PY
