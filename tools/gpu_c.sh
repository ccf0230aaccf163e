#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/c_stage.jsonl
python tools/stage_times.py > gpurun_out/c_stage.jsonl 2>gpurun_out/c_stage.err
for v in bb128x4 bb128x5; do
  KAMINO_B200_LIB=$PWD/kamino_b200/_lib/libkamino_b200_$v.so python tools/stage_times.py >> gpurun_out/c_stage.jsonl 2>>gpurun_out/c_stage.err
done
cat gpurun_out/c_stage.jsonl
python tools/stage_times.py 2 > gpurun_out/c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bloom_partition|bloom_bucket|cms_partition|cms_apply|frame_scatter' -s 15 -c 5 -o gpurun_out/r02_prof_c python tools/stage_times.py 2 > gpurun_out/c_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/c_ncu.log
ls -la gpurun_out/
