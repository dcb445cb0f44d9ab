#!/bin/bash
# Developer helper: resident C3 runs over the device feeder's concurrency knobs (RTJX_FEED_SLOTS, RTJX_LANES_SMEM_PAD).
run() { echo "== $*"; env "$@" RTJX_TRACE=1 timeout 600 python tools/prof_e2e.py 100000000 0 3 c3 resident 2>&1 | grep -v "^{" | grep -E "^rep 2|device feed" | tail -2 | cut -c1-260; }
run RTJX_FEED_SLOTS=10
run RTJX_FEED_SLOTS=6
run RTJX_FEED_SLOTS=8
run RTJX_FEED_SLOTS=14
run RTJX_FEED_SLOTS=16
run RTJX_FEED_SLOTS=10 RTJX_LANES_SMEM_PAD=4096
run RTJX_FEED_SLOTS=14 RTJX_LANES_SMEM_PAD=4096
run RTJX_FEED_SLOTS=10 RTJX_LANES_SMEM_PAD=10000
run RTJX_FEED_SLOTS=16 RTJX_GROUP_MB=128
