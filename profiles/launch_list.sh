#!/bin/bash
# launch list only (fast): device time and DRAM bytes of every launch.   usage: bash profiles/launch_list.sh <tag> [extra bench args]
# read with: python profiles/parse_launches.py gpurun_out/<tag>_launches.csv
TAG=${1:-r1}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-ref-gpu "$@" > gpurun_out/${TAG}_launches.log 2>&1
