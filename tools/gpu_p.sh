#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest_full.log; tail -4 gpurun_out/p_pytest_full.log
python tools/e2e_chunks.py > gpurun_out/p_e2e_chunks.jsonl 2> gpurun_out/p_e2e.err; cat gpurun_out/p_e2e_chunks.jsonl
timeout 900 python bench.py --no-strong > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/p_bench_n1.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "dram_frac", d["roofline"]["dram_fraction"])
print("pass2", json.dumps(d["pass2"])); print("nj", json.dumps(d["nj"]))
PY
