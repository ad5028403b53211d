#!/bin/bash
# scoring-kernel probe, both paths, deep shape slice (262144 reads x 1000 alleles x 300 kept sets)
for m in ${MODES:-1 0}; do
  echo "packed=$m"
  GK_PACKED=$m timeout 300 python tools/probe_kernels.py 262144 1000 2 2>&1 | grep -E "gk_score|kept" | tail -n 3
done
