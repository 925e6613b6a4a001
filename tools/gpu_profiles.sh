# round-end evidence: bench lines, ncu launch list and --set full captures of the top kernels (each after the same
# command has exited 0 without ncu)
set -x
python bench.py --steps 5 --warmup 3 2>gpurun_out/r1_bench_human.err | grep '^{' > gpurun_out/r1_bench_human.json; echo rc=$?
python bench.py --genome-bp 46000000 --steps 5 --warmup 3 2>gpurun_out/r1_bench_chr21.err | grep '^{' > gpurun_out/r1_bench_chr21.json; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/r1_bench_reference.err | grep '^{' > gpurun_out/r1_bench_reference.json; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:"probe_tm_kernel|collect_big_kernel|sw_pair_passes_kernel|sw_finish_band_kernel|minhash_warp_kernel|best_window_kernel" -s 30 -c 16 -o gpurun_out/r1_top_kernels -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_top.log 2>&1; echo rc=$?
ls -la gpurun_out/r1_*
