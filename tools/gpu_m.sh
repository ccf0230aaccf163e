#!/bin/bash
# named configurations at full size through the product CLI, beside the oracle CLI
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 2400 python tools/config_cli.py --species 200 --kind eukaryotic --devices "0" --oracle --seed-offset 3 > gpurun_out/m_c3_cli.json 2> gpurun_out/m_c3_cli.err
echo "c3 rc=$?"; cut -c1-2500 gpurun_out/m_c3_cli.json; tail -3 gpurun_out/m_c3_cli.err
