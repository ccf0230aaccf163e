#!/bin/bash
# round 2, first GPU call: parity of the new count-min partition (fixed slots + TMA stores) and a first bench line
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3diff or synthetic_small or edge or rare or saturation or full_size_c1 or staged" > gpurun_out/a_pytest1.log 2>&1
echo "pytest1 rc=$?" >> gpurun_out/a_pytest1.log
tail -5 gpurun_out/a_pytest1.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-strong > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"
cat gpurun_out/a_bench.json | head -c 6000
tail -5 gpurun_out/a_bench.err
