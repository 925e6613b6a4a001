set -x
BENCH_TRACE=1 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_final.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "text", d["e2e_text"]["value"], "fastq", d["e2e_text"]["from_fastq_text"]["value"])
print({k:v for k,v in d.items() if k.startswith("parity_ok") or k=="parity_checked"}, d["cpu_baseline"]["value"], d["roofline"]["frac"], d["clocks"])
PY
