#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/exchange_time.py > gpurun_out/j_xt_n$N.jsonl 2> gpurun_out/j_xt.err
grep rank gpurun_out/j_xt_n$N.jsonl
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu > gpurun_out/j_pytest_n$N.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/j_pytest_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/j_bench_n$N.json 2> gpurun_out/j_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/j_bench_n$N.json"))
print("weak ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("strong", json.dumps(d["strong_c2"]))
PY
