set -x
HRM_COLLECT_DEBUG=1 timeout 900 python tools/collect_stats.py 4000000 3100000000 2>&1 | grep -E "filter|rror|collectdiag" | tail -8
