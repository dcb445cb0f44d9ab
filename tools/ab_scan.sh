#!/bin/bash
# developer A/B harness for the cigar_scan variants (run under gpurun): ab_scan.sh "<variant> <debug>" ...
for cfg in "$@"; do
  set -- $cfg
  echo "== variant $1 debug $2"
  RTJX_SCAN_VARIANT=$1 RTJX_SCAN_DEBUG=$2 timeout 60 python tools/prof_step.py 10000000 6 2>&1 | grep -E "^step [45]|scan_ms" | sed -e 's/.*scan_ms/scan_ms/' | cut -c1-100
done
