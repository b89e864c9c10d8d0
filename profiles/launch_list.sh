#!/bin/bash
# launch list only (fast). usage: bash profiles/launch_list.sh <tag> [extra bench args]
TAG=${1:-r1}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_launches.log 2>&1
