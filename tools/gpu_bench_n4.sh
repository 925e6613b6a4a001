set -x
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo rc=$?
tail -c 600 gpurun_out/bench_n4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n4.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("text", d.get("e2e_text",{}).get("value"), d.get("e2e_text",{}).get("from_fastq_text",{}).get("value"))
print("part", json.dumps(d.get("partitioned_segment"))[:700])
PY
