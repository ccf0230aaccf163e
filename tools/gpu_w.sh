#!/bin/bash
# launch list (gpu__time_duration) of one warm pass-1 step + ncu --set full of the kernels matching $1
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_launches.csv python tools/stage_times.py 1 > gpurun_out/w_ncu1.log 2>&1
python tools/agg_launches.py gpurun_out/w_launches.csv 2>&1 | head -30
if [ -n "$1" ]; then bash tools/gpu_v.sh "$1" ${2:-2} ${3:-1}; fi
