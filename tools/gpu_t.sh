#!/bin/bash
# full GPU suite + bench N=1 (+ reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest_full.log; tail -4 gpurun_out/t_pytest_full.log
timeout 900 python bench.py > gpurun_out/t_bench_n1.json 2> gpurun_out/t_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/t_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/t_bench_n1.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "dram_frac", d["roofline"]["dram_fraction"])
print("kernels", [(k["kernel"][:22], round(k["ms"],3)) for k in d["roofline"]["kernels"]])
print("pass2", json.dumps(d["pass2"])); print("nj", json.dumps(d["nj"])); print("strong", json.dumps(d["strong_c2"])); print("cpu", json.dumps(d["cpu_baseline"]))
PY
