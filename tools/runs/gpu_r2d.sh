set -x
ncu --set full --clock-control none --import-source on -k regex:collect_dup -c 4 -o gpurun_out/collect_dup_r2 -f python tools/collect_stats.py 1000000 3100000000 > gpurun_out/ncu_r2d.log 2>&1; echo rc=$?
ncu -i gpurun_out/collect_dup_r2.ncu-rep --page raw --csv > gpurun_out/collect_dup_r2_raw.csv 2>/dev/null
ncu -i gpurun_out/collect_dup_r2.ncu-rep --page source --csv --print-source sass > gpurun_out/collect_dup_r2_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
rm -f gpurun_out/collect_dup_r2.ncu-rep
