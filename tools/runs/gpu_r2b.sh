set -x
python tools/collect_stats.py 1000000 3100000000 HRM_COLLECT_BLOOM_WORDS=1024 HRM_COLLECT_BLOOM_WORDS=4096 HRM_COLLECT_BLOOM_WORDS=8192 HRM_COLLECT_RANGES=1 2>&1 | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --reads 1000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2b.log 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/launches_r2b.csv | head -20
timeout 900 python -m pytest tests/test_gpu_store.py -x -q 2>&1 | tail -15
