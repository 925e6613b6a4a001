# usage: gpu_ncu_kernel.sh <kernel regex> <skip> <count> <tag>
set -x
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$4.log 2>&1; echo rc=$?
ls -la gpurun_out/$4.ncu-rep
