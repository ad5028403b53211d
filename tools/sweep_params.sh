#!/bin/bash
# kernel-time sweep over the work-item granularities (cfg5 cohort, 96 samples)
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_gpu_host.py 96 2>&1 | grep -E "gk_write_p|gk_score  |gk_rescore_count|kernel sum"; }
run GK_MIN_SCORE_ITEMS=1184
run GK_MIN_SCORE_ITEMS=1776
run GK_MIN_SCORE_ITEMS=3552
run GK_MIN_SCORE_ITEMS=600
run GK_P_CHUNK=1024
run GK_P_CHUNK=4096
run GK_COUNT_CHUNK=4096
run GK_COUNT_CHUNK=8192
