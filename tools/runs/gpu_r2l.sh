set -x
for cfgs in "4 0" "2 0" "2 6"; do
set -- $cfgs
HRM_K7A_BLOCKS_PER_SM=$1 HRM_COLLECT_BLOCKS_PER_SM=$2 python bench.py --steps 6 --warmup 3 --check 0 --no-cpu-baseline --no-text > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2l.json") if l.startswith("{")][0])
print("K7A/COLLECT per SM = $cfgs: value", d["value"], "serial", d["value_one_batch_at_a_time"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["ms_per_step_one_batch_at_a_time"])
PY
done
