#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3diff or synthetic_small or edge or rare or saturation or full_size_c1 or staged or bloom" > gpurun_out/s_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/s_pytest1.log; tail -3 gpurun_out/s_pytest1.log
python tools/stage_times.py > gpurun_out/s_stage.jsonl 2>gpurun_out/s_stage.err; cat gpurun_out/s_stage.jsonl
