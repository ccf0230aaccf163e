#!/bin/bash
# 8-GPU box: multi-GPU parity at 2/4/8 ranks, the multi-device CLI, configs[2] through the CLI on 1 and 8 GPUs,
# exchange timing with NVLink byte counters, bench (weak + strong C2)
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/l_smi.txt; nproc >> gpurun_out/l_smi.txt
timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_e2e.py -x -q -m gpu -k "multi_device or multi_gpu or timeout or pair_counts_row" > gpurun_out/l_pytest_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest_n$N.log; tail -4 gpurun_out/l_pytest_n$N.log
DEVS=$(python -c "print(','.join(map(str, range($N))))")
timeout 900 python tools/config_cli.py --species 1000 --devices "0;$DEVS" > gpurun_out/l_c2_cli_n$N.json 2> gpurun_out/l_c2_cli.err
echo "c2 rc=$?"; cut -c1-3000 gpurun_out/l_c2_cli_n$N.json; tail -3 gpurun_out/l_c2_cli.err
nvidia-smi nvlink -gt d > gpurun_out/l_nvlink_before.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/exchange_time.py > gpurun_out/l_xt_n$N.jsonl 2> gpurun_out/l_xt.err
nvidia-smi nvlink -gt d > gpurun_out/l_nvlink_after.txt 2>&1
grep '"rank": 0' gpurun_out/l_xt_n$N.jsonl
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/l_bench_n$N.json 2> gpurun_out/l_bench_n$N.err
echo "bench rc=$?"; grep -E "checks passed|strong" gpurun_out/l_bench_n$N.err | head -4
python - <<PY
import json
d=json.load(open("gpurun_out/l_bench_n$N.json"))
print("weak ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("strong", json.dumps(d["strong_c2"]))
PY
