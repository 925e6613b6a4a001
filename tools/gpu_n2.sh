set -x
python -m pytest tests/test_gpu_partitioned.py -x -q 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29701 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench_n2.err | grep '^{' > gpurun_out/bench_n2.json; echo rc=$?
$TR --master-port 29702 bench.py --gpus 2 --steps 3 --warmup 3 --index partitioned 2>gpurun_out/bench_n2p.err | grep '^{' > gpurun_out/bench_n2p.json; echo rc=$?
tail -5 gpurun_out/bench_n2p.err
python - <<PY
import json
for f in ("bench_n2","bench_n2p"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["n_gpus"], d["value"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("stages_ms_per_step"), d["config"]["parallelism"], d.get("exchange"), d["config"]["index_device_bytes"], d.get("collect_ids_skipped_fraction"))
    except Exception as e:
        print(f, "failed", e)
PY
