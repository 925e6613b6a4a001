set -x
timeout 900 python -m pytest tests/test_gpu_store.py -x -q 2>&1 | tail -6
ncu --set full --clock-control none --import-source on -k regex:collect_dup -c 2 -o gpurun_out/collect_dup_r2b -f python tools/collect_stats.py 1000000 3100000000 > gpurun_out/ncu_r2k.log 2>&1; echo rc=$?
ncu -i gpurun_out/collect_dup_r2b.ncu-rep --page raw --csv > gpurun_out/collect_dup_r2b_raw.csv 2>/dev/null
ncu -i gpurun_out/collect_dup_r2b.ncu-rep --page source --csv --print-source sass > gpurun_out/collect_dup_r2b_source.csv 2>/dev/null
rm -f gpurun_out/collect_dup_r2b.ncu-rep
