set -x
cap() { # name regex skip count
ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/r1_$1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1; echo rc=$?
}
cap probe_tm probe_tm_kernel 8 2
cap collect_big collect_big_kernel 8 1
cap sw_pair sw_pair_passes_kernel 4 1
cap sw_band sw_finish_band_kernel 32 7
cap shd_minhash "best_window_kernel|collect_small_kernel|transpose_sigs_kernel|totals_tm_kernel" 24 8
ls -la gpurun_out/r1_*.ncu-rep
