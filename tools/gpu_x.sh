#!/bin/bash
# parity subset + launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3diff or synthetic_small or edge or rare or saturation or full_size_c1 or staged or bloom or frame or fasta" > gpurun_out/u_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/u_pytest1.log; tail -5 gpurun_out/u_pytest1.log
timeout 120 python tools/stage_times.py > gpurun_out/u_stage.jsonl 2>gpurun_out/u_stage.err; cat gpurun_out/u_stage.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_launches.csv python tools/stage_times.py 1 > gpurun_out/w_ncu1.log 2>&1
python tools/agg_launches.py gpurun_out/w_launches.csv 2>&1 | head -8
