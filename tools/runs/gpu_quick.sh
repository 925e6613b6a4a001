# quick GPU check: parity tests + one bench line (+ optional launch list)
set -x
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo rc=$?
tail -c 400 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"])
PY
if [ -n "$2" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1; echo rc=$?
fi
