#!/bin/bash
# Pass pipelining A/B on one GPU: 12 samples (the share of one rank of the 8-GPU cohort run) and the
# full 96-sample cohort, against the number of passes in flight.  Output: gpurun_out/pipe_*.json
set -x
mkdir -p gpurun_out
common="--no-deep --no-cpu-baseline --steps 40"
for d in ${DEPTHS:-3 4}; do
timeout 200 python bench.py $common --samples 12 --pipeline-depth $d > gpurun_out/pipe_s12_d$d.json 2> gpurun_out/pipe_s12_d$d.err
done
timeout 300 python bench.py $common --steps 20 --pipeline-depth 3 > gpurun_out/pipe_s96_d3.json 2> gpurun_out/pipe_s96_d3.err
echo done
