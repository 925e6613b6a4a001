set -x
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo rc=$?
tail -c 300 gpurun_out/bench_r2m.err
# launch list of the hot kernels of the same command (filtered: the index build's thousands of launches are not profiled)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"collect_|sw_|probe_tm|minhash_warp|best_window|sam_|pack_rows|merge_pass|record_header|scatter_align" -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --check 0 > gpurun_out/ncu_r2m.log 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/r2_launches.csv | head -24
# full capture of the dominant HBM-bound kernel at the bench's batch size (4 M reads per launch)
ncu --set full --clock-control none --import-source on -k regex:collect_dup -c 2 -o gpurun_out/collect_dup_4m -f python tools/collect_stats.py 4000000 3100000000 > gpurun_out/ncu_r2m2.log 2>&1; echo rc=$?
ncu -i gpurun_out/collect_dup_4m.ncu-rep --page raw --csv > gpurun_out/r2_collect_dup_4m_raw.csv 2>/dev/null
rm -f gpurun_out/collect_dup_4m.ncu-rep
