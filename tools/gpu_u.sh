#!/bin/bash
# A/B of tuning builds: parity subset with the default library, then per-kernel stage times of every variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3diff or synthetic_small or edge or rare or saturation or full_size_c1 or staged or bloom or frame or fasta" > gpurun_out/u_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/u_pytest1.log; tail -15 gpurun_out/u_pytest1.log
: > gpurun_out/u_stage.jsonl
timeout 120 python tools/stage_times.py >> gpurun_out/u_stage.jsonl 2>gpurun_out/u_stage.err
for v in "$@"; do
  KAMINO_B200_LIB=$PWD/kamino_b200/_lib/libkamino_b200_$v.so timeout 120 python tools/stage_times.py >> gpurun_out/u_stage.jsonl 2>>gpurun_out/u_stage.err
done
cat gpurun_out/u_stage.jsonl; tail -5 gpurun_out/u_stage.err
