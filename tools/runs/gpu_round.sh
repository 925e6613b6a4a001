set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; echo rc=$?
tail -c 1500 gpurun_out/bench_v9.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v9.json 2> gpurun_out/bench_ref_v9.err; echo rc=$?
cat gpurun_out/bench_ref_v9.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1_v9.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v9.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:probe_count_kernel -s 4 -c 2 -o gpurun_out/probe_v9 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_v9.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:sw_pair_passes_kernel -s 1 -c 1 -o gpurun_out/swpair_v9 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_v9.log 2>&1; echo rc=$?
ls -la gpurun_out
