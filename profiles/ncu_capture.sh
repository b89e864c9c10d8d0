#!/bin/bash
# Run on the GPU box (under gpurun): launch list, DRAM bytes of every kernel of one multiply, and full captures
# of the dominant numeric and symbolic kernels.   usage: bash profiles/ncu_capture.sh <tag>   -> gpurun_out/<tag>_*
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-ref-gpu"
# 1) every launch with its device time and DRAM bytes (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
# 2) mapped numeric CTA kernels of the second multiply (5 shapes) + the lane-group kernel of the 512 class
ncu --set full --clock-control none --import-source on -k regex:"k_map_rows" -s 13 -c 6 \
    -o $OUT/${TAG}_numeric $BENCH > $OUT/${TAG}_numeric.log 2>&1
# 3) mapped symbolic rank kernels of the second multiply (5 shapes)
ncu --set full --clock-control none --import-source on -k regex:"k_rank_rows" -s 5 -c 5 \
    -o $OUT/${TAG}_symbolic $BENCH > $OUT/${TAG}_symbolic.log 2>&1
ls -la $OUT
