set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for mode in classify b1; do
  if [ $mode = b1 ]; then export HRM_SW_B1=1; else unset HRM_SW_B1; fi
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$mode.json 2> gpurun_out/bench_ab_$mode.err; echo rc=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab_$mode.json"))
print("$mode", d["value"], d["e2e"]["value"], d["stages_ms_per_step"]["verify"])
PY
done
unset HRM_SW_B1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_ab.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ab.log 2>&1; echo rc=$?
