#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/e2e_chunks.py > gpurun_out/o_e2e_chunks.jsonl 2> gpurun_out/o_e2e.err; cat gpurun_out/o_e2e_chunks.jsonl
python bench.py --steps 1 --warmup 3 --profile > gpurun_out/o_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --profile > gpurun_out/o_ncu1.log 2>&1
echo "ncu1 rc=$?"
python tools/stage_times.py 2 > gpurun_out/o_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bloom_partition|bloom_bucket|cms_partition|cms_apply|frame_scatter|frame_count' -s 18 -c 6 -o gpurun_out/r02_prof_final python tools/stage_times.py 2 > gpurun_out/o_ncu2.log 2>&1
echo "ncu2 rc=$?"; ls -la gpurun_out | grep -E "r02_prof_final|r02_launches"
