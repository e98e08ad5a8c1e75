#!/bin/bash
# One gpurun call of the development loop: GPU tests, bench lines, logs into gpurun_out/ (scratch).
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag> [tests|bench|all]'
tag=${1:-x}
what=${2:-all}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/${tag}_smi.txt 2>&1
if [ "$what" = "tests" ] || [ "$what" = "all" ]; then
  timeout -s KILL 1200 python -m pytest tests -q -m gpu --maxfail=15 -p no:cacheprovider --timeout 600 \
      > $out/${tag}_tests.log 2>&1
  echo "tests exit $?" >> $out/${tag}_tests.log
  tail -n 40 $out/${tag}_tests.log
fi
if [ "$what" = "bench" ] || [ "$what" = "all" ]; then
  timeout -s KILL 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
  echo "bench exit $?"; tail -c 1500 $out/${tag}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("value", "ms_per_step", "roofline", "roofline_hessian", "adaptive_solve", "e2e", "e2e_solve", "configs", "clocks")}
    print(json.dumps(keep, indent=1)[:6000])
except Exception as e:
    print("no bench line:", e)
PY
fi
