set -x
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mapper.py -x -q -k "unique_by_count or baseline_configs or ingest" 2>&1 | tail -15
