#!/bin/bash
# Evidence run on one B200 (everything lands in gpurun_out/; the digests are copied to profiles/ by hand):
#   - the two microbenchmarks behind the roofline denominators (non-tensor pipe mix, bulk copy),
#   - the ncu launch list of the default bench command (kernel shares of a step),
#   - ncu --set full on every gk_* kernel of one cohort pass, digested HERE into text (summary per launch,
#     stall reasons and hottest SASS lines of the scoring kernel, DRAM bytes per launch): the report itself
#     exceeds what gpurun brings back and is deleted.
# A number printed by a run under ncu is never a bench value.
set -u
R=${1:-r02}
OUT=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mixpipe tools/micro/mixpipe.cu 2>/dev/null && /tmp/mixpipe > $OUT/${R}_mixpipe.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulkcopy tools/micro/bulkcopy.cu 2>/dev/null && /tmp/bulkcopy > $OUT/${R}_bulkcopy.txt 2>&1
COMMON="--no-deep --no-cpu-baseline --no-cold --no-host"
python bench.py --steps 2 --warmup 3 $COMMON > $OUT/${R}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 $COMMON > $OUT/${R}_launches_run.log 2>&1
GK_GRAPH=0 ncu --set full --clock-control none -k regex:gk_ --launch-skip 54 -c 27 -f -o /tmp/${R}_pass \
    python bench.py --steps 3 --warmup 3 $COMMON --pipeline-depth 1 > $OUT/${R}_pass_run.log 2>&1
python tools/ncu_summary.py /tmp/${R}_pass.ncu-rep > $OUT/${R}_pass_ncu_summary.txt 2>&1
cp profiles/traffic.json $OUT/${R}_traffic.json
python tools/ncu_traffic.py /tmp/${R}_pass.ncu-rep $OUT/${R}_traffic.json > $OUT/${R}_traffic_digest.txt 2>&1
ls -la $OUT /tmp/${R}_pass.ncu-rep
