#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3diff or synthetic_small or edge or rare or saturation or full_size_c1 or staged or bloom" > gpurun_out/n_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/n_pytest1.log; tail -3 gpurun_out/n_pytest1.log
python tools/stage_times.py > gpurun_out/n_stage.jsonl 2>gpurun_out/n_stage.err; cat gpurun_out/n_stage.jsonl
df -h /tmp | tail -1
timeout 3000 python tools/config_cli.py --species 5000 --nj --devices "0" --oracle --seed-offset 4 > gpurun_out/n_c4_cli.json 2> gpurun_out/n_c4_cli.err
echo "c4 rc=$?"; cut -c1-3000 gpurun_out/n_c4_cli.json; tail -3 gpurun_out/n_c4_cli.err
