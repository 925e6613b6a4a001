set -x
timeout 900 python -m pytest tests/test_gpu_partitioned.py -x -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2h_n2.json 2> gpurun_out/bench_r2h_n2.err; echo rc=$?
tail -c 800 gpurun_out/bench_r2h_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2h_n2.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "text", (d.get("e2e_text") or {}).get("value"))
print("stages", d["stages_ms_per_step"])
print("part", d.get("partitioned_segment"))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --index partitioned --no-text > gpurun_out/bench_r2h_n2_part.json 2> gpurun_out/bench_r2h_n2_part.err; echo rc=$?
tail -c 500 gpurun_out/bench_r2h_n2_part.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2h_n2_part.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print("stages", d["stages_ms_per_step"], d.get("exchange"))
PY
