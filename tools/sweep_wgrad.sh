#!/bin/bash
# In-situ sweep of the weight-gradient N-chunking (GEECO_TC_WGRAD_NSUB) and the k-blocks-per-split floor
# (GEECO_TC_WGRAD_MINKB): one timeline of the real batch-64 step per setting, weight-gradient kernels + their reduces.
for cfg in "0 8" "1 8" "2 8" "3 8" "4 8" "0 4" "2 4" "4 4"; do
  set -- $cfg
  echo "== nsub $1 minkb $2"
  GEECO_TC_WGRAD_NSUB=$1 GEECO_TC_WGRAD_MINKB=$2 timeout 120 python tools/timeline.py 2>&1 | grep -E "tc_wgrad_kernel|wgrad_reduce|span" | awk '{printf "%s %s | ", $2, substr($5,1,28)} END {print ""}'
done
