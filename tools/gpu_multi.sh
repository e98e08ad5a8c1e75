#!/bin/bash
# Multi-GPU evidence in one call:  gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh <tag> N'
tag=${1:-mg}
n=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_gpus.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu -p no:cacheprovider > $out/${tag}_mgtest.log 2>&1
echo "mg test exit $?"; tail -5 $out/${tag}_mgtest.log; tail -3 $out/mg_worker.log
port=$((29700 + RANDOM % 200))
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 100 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
echo "bench exit $?"; tail -c 600 $out/${tag}_bench_n${n}.err
for cfg in c4 c5; do
  port=$((29700 + RANDOM % 200))
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --config $cfg > $out/${tag}_${cfg}_n${n}.json 2> $out/${tag}_${cfg}_n${n}.err
  echo "$cfg exit $?"
done
python - <<PY
import json
for name in ("bench", "c4", "c5"):
    try:
        d = json.loads(open("$out/${tag}_%s_n$n.json" % name).read().strip().splitlines()[-1])
        if name == "bench":
            print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "parity_multi_rank", "roofline")}, indent=1)[:2500])
            print(json.dumps(d.get("configs", {}).get("c4"), indent=1)[:1500])
        else:
            print(json.dumps(d["config"], indent=1)[:1800])
    except Exception as e:
        print(name, "no line:", e)
PY
