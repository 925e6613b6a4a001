set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/sweep_kw.py --reads 200000 --out gpurun_out/sweep_kw_n2.json > gpurun_out/sweep_n2.log 2>&1; echo rc=$?
grep "^{" gpurun_out/sweep_n2.log | cut -c1-700
tail -5 gpurun_out/sweep_n2.log | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 2 --warmup 3 --index partitioned --no-text > gpurun_out/bench_r2i_n2_part.json 2> gpurun_out/bench_r2i_n2_part.err; echo rc=$?
tail -c 300 gpurun_out/bench_r2i_n2_part.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2i_n2_part.json") if l.startswith("{")][0])
print("value", d["value"], "e2e", d["e2e"]["value"])
print("stages", d["stages_ms_per_step"], d.get("exchange"))
PY
