# usage: gpu_ncu_any.sh <kernel regex> <skip> <count> <tag> <bench args...>
set -x
K=$1; S=$2; C=$3; TAG=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/$TAG -f python bench.py "$@" > gpurun_out/ncu_$TAG.log 2>&1; echo rc=$?
ls -la gpurun_out/$TAG.ncu-rep
