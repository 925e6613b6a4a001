set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sam_|scan" -c 400 --csv --log-file gpurun_out/launches_sam.csv python bench.py --reads 1000000 --steps 2 --warmup 3 --no-cpu-baseline --check 0 --no-fastq > gpurun_out/ncu_sam.log 2>&1; echo rc=$?
python tools/launch_summary.py gpurun_out/launches_sam.csv | head -12
