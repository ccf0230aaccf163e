#!/bin/bash
# end-of-round evidence: full GPU suite, bench N=1 (+ reference arm), launch list of the bench under ncu, ncu --set full of
# one warm launch of every hot kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest_full.log; tail -4 gpurun_out/f_pytest_full.log
timeout 900 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/f_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
echo "ref rc=$?"; cat gpurun_out/f_bench_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv python bench.py --profile --steps 1 --warmup 3 > gpurun_out/f_ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/agg_launches.py gpurun_out/f_launches.csv 2>&1 | head -40
rm -f gpurun_out/f_prof.ncu-rep
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"frame_kernel|bloom_partition_kernel|bloom_bucket_kernel|cms_partition_kernel|cms_apply_kernel" --launch-skip 10 --launch-count 5 -f -o gpurun_out/f_prof python tools/stage_times.py 1 > gpurun_out/f_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/f_ncu_full.log
