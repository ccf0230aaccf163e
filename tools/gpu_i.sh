#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
rm -f gpurun_out/i_xt.jsonl
for v in "XT_GAP=none" "XT_GAP=sleep" "XT_GAP=barrier" "KMN_EXCHANGE_NOABORT=1"; do
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/exchange_time.py >> gpurun_out/i_xt.jsonl 2> gpurun_out/i_xt.err
done
grep rank gpurun_out/i_xt.jsonl
