#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest_full.log; tail -4 gpurun_out/t_pytest_full.log
timeout 1500 python tools/config_cli.py --species 1000 --devices "0" --oracle > gpurun_out/t_c2_cli_1gpu.json 2> gpurun_out/t_c2_cli.err
echo "c2 rc=$?"; cut -c1-1800 gpurun_out/t_c2_cli_1gpu.json; tail -3 gpurun_out/t_c2_cli.err
timeout 900 python bench.py > gpurun_out/t_bench_n1.json 2> gpurun_out/t_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/t_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err
echo "ref rc=$?"; cat gpurun_out/t_bench_ref.json | cut -c1-900
python - <<PY
import json
d=json.load(open("gpurun_out/t_bench_n1.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "dram_frac", d["roofline"]["dram_fraction"])
print("kernels", [(k["kernel"][:22], round(k["ms"],3)) for k in d["roofline"]["kernels"]])
print("pass2", json.dumps(d["pass2"])); print("nj", json.dumps(d["nj"])); print("strong", json.dumps(d["strong_c2"]))
PY
