set -x
ncu --set full --clock-control none -k regex:probe_tm_kernel -s 8 -c 2 -o /tmp/r1_probe4m -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_probe4m.log 2>&1; echo rc=$?
ncu -i /tmp/r1_probe4m.ncu-rep --page raw --csv > gpurun_out/r1_probe_tm_4m_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r1_probe_tm_4m_raw.csv")))
h=rows[0]; u=rows[1]
for k in ("gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","sm__warps_active.avg.pct_of_peak_sustained_active"):
    i=h.index(k); print(k,u[i],[r[i] for r in rows[2:]])
PY
