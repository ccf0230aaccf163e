#!/bin/bash
# N-GPU checks: exchange parity + timeout, in-process multi-device CLI, bench with the strong-scaling C2 job
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/f_smi_$N.txt
timeout 1200 python -m pytest tests/test_multigpu.py tests/test_gpu_e2e.py -x -q -m gpu -k "multi_device or multi_gpu or timeout or pair_counts_row" > gpurun_out/f_pytest_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest_$N.log
tail -8 gpurun_out/f_pytest_$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/f_bench_n$N.json 2> gpurun_out/f_bench_n$N.err
echo "bench rc=$?"
cat gpurun_out/f_bench_n$N.json | head -c 5000
grep -E "rank 0|checks|strong|Error|error" gpurun_out/f_bench_n$N.err | tail -12
