#!/bin/bash
# A/B: tools/stage_times.py on the default library and on the tuning builds named as arguments (twice, interleaved)
mkdir -p gpurun_out
out=gpurun_out/${AB_OUT:-ab}.jsonl; : > $out
for rep in 1 2; do
for l in "" "$@"; do
  if [ -n "$l" ]; then export KAMINO_B200_LIB=$PWD/kamino_b200/_lib/libkamino_b200_$l.so; else unset KAMINO_B200_LIB; fi
  python tools/stage_times.py 10 >> $out
done; done
python - <<PY
import json
for l in open("$out"):
    d=json.loads(l); m=d["ms"]
    print(d["lib"].split("libkamino_b200")[-1][:12].ljust(12), m["frame"], m["bloom_partition_kernel"], m["bloom_bucket_kernel"], m["cms_partition_kernel"], m["cms_apply_kernel"], "sum", m["pass1_sum"], d["table_hash"], d["new_kmers"])
PY
