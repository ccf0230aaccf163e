#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/v_prof.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"bloom_bucket_kernel|frame_kernel|bloom_partition_kernel" --launch-skip 6 --launch-count 3 -f -o gpurun_out/v_prof python tools/stage_times.py 1 > gpurun_out/v_ncu.log 2>&1
tail -5 gpurun_out/v_ncu.log
