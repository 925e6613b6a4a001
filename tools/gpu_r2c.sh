set -x
HRM_COLLECT_DEBUG=1 python tools/collect_stats.py 1000000 3100000000 HRM_COLLECT_BLOOM_BLOCK_WORDS=32768,HRM_COLLECT_BLOCK_XSLOTS=8192 2>&1 | grep -v "^Traceback\|AttributeError\|File \|Exception ignored" | tail -5
timeout 1200 python -m pytest tests/test_gpu_mapper.py tests/test_gpu_store.py -x -q -k "fused_collection or dump" 2>&1 | tail -8
