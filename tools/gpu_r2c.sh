set -x
HRM_COLLECT_DEBUG=1 python tools/collect_stats.py 1000000 3100000000 HRM_COLLECT_BLOOM_WORDS=2048 HRM_COLLECT_XSLOTS=512 HRM_COLLECT_BLOOM_WORDS=512 HRM_COLLECT_RANGES=1 2>&1 | grep -v "^Traceback\|AttributeError\|File \|Exception ignored" | grep -v "^collect" | tail -8
timeout 1200 python -m pytest tests/test_gpu_mapper.py -x -q -k "fused_collection" 2>&1 | tail -8
