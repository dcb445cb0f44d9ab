#!/bin/bash
# developer A/B harness for cigar_scan (run under gpurun): ab_scan.sh "ENV=VAL ENV=VAL" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 60 python tools/prof_step.py 10000000 8 2>&1 | grep -E "^step [67]|scan_ms" | sed -e 's/.*scan_ms/scan_ms/' | cut -c1-100
done
