set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo rc=$?
tail -c 600 gpurun_out/bench_r2e.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2e.json"))
print("value", d["value"], "e2e", d["e2e"], "text", d.get("e2e_text"))
print("stages", d["stages_ms_per_step"])
print("roofline", {k:v for k,v in d["roofline"].items() if k!="note"})
print({k:v for k,v in d.items() if k.startswith("parity")})
print(d.get("cpu_baseline"))
PY
