#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3diff or synthetic_small or edge or rare or saturation or full_size_c1 or staged or bloom" > gpurun_out/b_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/b_pytest1.log
tail -5 gpurun_out/b_pytest1.log
python tools/stage_times.py > gpurun_out/b_stage.jsonl 2>gpurun_out/b_stage.err
for v in bb256x4 bb256x3 bb1024x1; do
  KAMINO_B200_LIB=$PWD/kamino_b200/_lib/libkamino_b200_$v.so python tools/stage_times.py >> gpurun_out/b_stage.jsonl 2>>gpurun_out/b_stage.err
done
cat gpurun_out/b_stage.jsonl
tail -3 gpurun_out/b_stage.err
