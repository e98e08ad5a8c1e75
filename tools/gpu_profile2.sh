#!/bin/bash
# launch lists (kernel shares) of the C2 / C4 config lines:  gpurun -- 'bash tools/gpu_profile2.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
for cfg in c2 c4; do
  timeout -s KILL 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 3000 --csv \
      --log-file $out/launches_${cfg}_${tag}.csv python bench.py --config $cfg > $out/bench_${cfg}_under_ncu_${tag}.log 2>&1
done
ls -la $out/launches_c*_${tag}.csv
