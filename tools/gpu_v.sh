#!/bin/bash
# ncu --set full of one warm launch of the kernels matching $1 (regex), report -> gpurun_out/v_prof.ncu-rep
mkdir -p gpurun_out
rm -f gpurun_out/v_prof.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$1" --launch-skip ${2:-2} --launch-count ${3:-1} -f -o gpurun_out/v_prof python tools/stage_times.py 1 > gpurun_out/v_ncu.log 2>&1
tail -4 gpurun_out/v_ncu.log
