set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/sweep_kw.py --reads 200000 --out gpurun_out/sweep_kw_n8.json > gpurun_out/sweep_n8.log 2>&1; echo rc=$?
grep "^{" gpurun_out/sweep_n8.log | cut -c1-500
tail -3 gpurun_out/sweep_n8.log | cut -c1-300
