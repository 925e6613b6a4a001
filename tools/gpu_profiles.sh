# round-end evidence: bench lines, ncu launch list and --set full captures of the top kernels (each after the same
# command has exited 0 without ncu).  Reports stay on the box; their raw pages come back as CSV.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu.txt; cat gpurun_out/r1_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 2>gpurun_out/r1_bench_human.err | grep '^{' > gpurun_out/r1_bench_human.json; echo rc=$?
python bench.py --genome-bp 46000000 --steps 5 --warmup 3 2>gpurun_out/r1_bench_chr21.err | grep '^{' > gpurun_out/r1_bench_chr21.json; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/r1_bench_reference.err | grep '^{' > gpurun_out/r1_bench_reference.json; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
cap() { # name regex skip count
ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/r1_$1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1; echo rc=$?
ncu -i /tmp/r1_$1.ncu-rep --page raw --csv > gpurun_out/r1_$1_raw.csv 2>/dev/null
}
cap probe_tm probe_tm_kernel 8 2
cap collect_warp collect_warp_kernel 8 1
ncu -i /tmp/r1_collect_warp.ncu-rep --page source --print-source sass --csv > gpurun_out/r1_collect_warp_sass.csv 2>/dev/null
cap sw_pair sw_pair_passes_kernel 4 1
cap sw_band sw_finish_band_kernel 32 8
cap k2_k5 "best_window_kernel|minhash_warp_kernel" 66 4
du -sh gpurun_out; ls -la gpurun_out
