#!/bin/bash
# kernel-time sweep over the work-item granularities (cfg5 cohort, 96 samples)
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_gpu_host.py 96 2>&1 | grep -E "gk_write_p|gk_score  |gk_rescore_count|kernel sum"; }
run GK_SCORE_CHUNK=8192
run GK_SCORE_CHUNK=4096
run GK_SCORE_CHUNK=2048
run GK_SCORE_CHUNK=1024
run GK_P_CHUNK=1024
run GK_P_CHUNK=4096
run GK_COUNT_CHUNK=4096
run GK_COUNT_CHUNK=8192
