set -x
timeout 900 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu_n2.txt; cat gpurun_out/r2_pytest_gpu_n2.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_validate_n2.json 2> gpurun_out/bench_validate_n2.err; echo rc=$?
tail -c 400 gpurun_out/bench_validate_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_validate_n2.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("text", d.get("e2e_text",{}).get("value"))
print("part", json.dumps(d.get("partitioned_segment"))[:900])
PY
