#!/bin/bash
# Developer helper: resident C3 runs over the device feeder's group size / slot count (RTJX_GROUP_MB, RTJX_FEED_SLOTS).
run() { echo "== $*"; env "$@" RTJX_TRACE=1 timeout 600 python tools/prof_e2e.py 100000000 0 3 c3 resident 2>&1 | grep -v "^{" | grep -E "^rep 2|device feed" | tail -2 | cut -c1-230; }
run RTJX_GROUP_MB=256 RTJX_FEED_SLOTS=10
run RTJX_GROUP_MB=512 RTJX_FEED_SLOTS=10
run RTJX_GROUP_MB=512 RTJX_FEED_SLOTS=6
run RTJX_GROUP_MB=384 RTJX_FEED_SLOTS=8
run RTJX_GROUP_MB=1024 RTJX_FEED_SLOTS=5
run RTJX_GROUP_MB=512 RTJX_FEED_SLOTS=10 RTJX_FIRST_GROUP_MB=32
