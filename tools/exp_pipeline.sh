#!/bin/bash
# Pass pipelining A/B on one GPU: 12 samples (the share of one rank of the 8-GPU cohort run) and the
# full 96-sample cohort, with one and two passes in flight.  Output: gpurun_out/pipe_*.json
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "pass_pipeline or two_batches or repeated_passes" > gpurun_out/pipe_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/pipe_tests.log
common="--no-deep --no-cpu-baseline --steps 40"
timeout 200 python bench.py $common --samples 12 --pipeline-depth 1 > gpurun_out/pipe_s12_d1.json 2> gpurun_out/pipe_s12_d1.err
timeout 200 python bench.py $common --samples 12 --pipeline-depth 2 > gpurun_out/pipe_s12_d2.json 2> gpurun_out/pipe_s12_d2.err
timeout 200 python bench.py $common --samples 12 --pipeline-depth 2 --parts 2 > gpurun_out/pipe_s12_d2p2.json 2> gpurun_out/pipe_s12_d2p2.err
timeout 300 python bench.py $common --steps 20 --pipeline-depth 2 > gpurun_out/pipe_s96_d2.json 2> gpurun_out/pipe_s96_d2.err
echo done
