set -x
HRM_COLLECT_DEBUG=1 timeout 900 python tools/collect_stats.py 4000000 3100000000 HRM_COLLECT_XSLOTS=1024 HRM_COLLECT_BLOCKS_PER_SM=10 2>&1 | grep -E "filter|rror" | tail -8
