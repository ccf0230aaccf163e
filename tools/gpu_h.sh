#!/bin/bash
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/exchange_time.py > gpurun_out/h_xt_n$N.jsonl 2> gpurun_out/h_xt_n$N.err
echo "rc=$?"; cat gpurun_out/h_xt_n$N.jsonl; tail -3 gpurun_out/h_xt_n$N.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_e2e.py -x -q -m gpu -k "group or synthetic_matches or golden" > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/h_pytest.log
