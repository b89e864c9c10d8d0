#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the dominant kernels.
# usage: bash profiles/ncu_capture.sh <tag>      outputs -> gpurun_out/<tag>_*
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
# 1) every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
# 2) dense (bitmap) kernels: skip the first multiply, capture symbolic + numeric of the second
ncu --set full --clock-control none --import-source on -k regex:k_dense_rows -s 2 -c 2 \
    -o $OUT/${TAG}_dense $BENCH > $OUT/${TAG}_dense.log 2>&1
# 3) CTA sort kernels of the second multiply (3 symbolic + 3 numeric)
ncu --set full --clock-control none --import-source on -k regex:k_sort_rows_cta -s 6 -c 6 \
    -o $OUT/${TAG}_sortcta $BENCH > $OUT/${TAG}_sortcta.log 2>&1
ls -la $OUT
