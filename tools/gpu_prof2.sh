#!/bin/bash
# ncu evidence beyond pass 1's five kernels: cms_apply_kernel merging into a live table, and the pass-2 kernels
mkdir -p gpurun_out
rm -f gpurun_out/g_apply_live.ncu-rep gpurun_out/g_pass2.ncu-rep
KMN_STAGE_ACCUMULATE=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"cms_apply_kernel" --launch-skip 3 --launch-count 1 -f -o gpurun_out/g_apply_live python tools/stage_times.py 1 > gpurun_out/g_apply_live.log 2>&1
echo "apply live rc=$?"
timeout 300 python tools/pass2_run.py 3 > gpurun_out/g_pass2.json 2> gpurun_out/g_pass2.err; echo "pass2 rc=$?"; cat gpurun_out/g_pass2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/g_pass2_launches.csv python tools/pass2_run.py 2 > gpurun_out/g_pass2_launch.log 2>&1
echo "pass2 launch list rc=$?"
python tools/agg_launches.py gpurun_out/g_pass2_launches.csv 2>&1 | head -40
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"occupancy_sweep8_kernel|segment_boundary_kernel|pair_segments_kernel|radix_scatter_kernel" --launch-skip 14 --launch-count 8 -f -o gpurun_out/g_pass2 python tools/pass2_run.py 2 > gpurun_out/g_pass2_full.log 2>&1
echo "pass2 full rc=$?"; tail -2 gpurun_out/g_pass2_full.log
