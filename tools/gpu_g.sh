#!/bin/bash
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/g_bench_n$N.json 2> gpurun_out/g_bench_n$N.err
echo "bench rc=$?"
cat gpurun_out/g_bench_n$N.json | head -c 6000
grep -E "rank 0|checks|strong|Error|error" gpurun_out/g_bench_n$N.err | tail -12
