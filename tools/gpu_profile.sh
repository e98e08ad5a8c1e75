#!/bin/bash
# ncu captures of the round-2 kernels (one GPU; results in gpurun_out/, summaries are made in the build container
# with tools/summarize_ncu.py).   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# 1. launch list of a short default bench run (kernel shares of the step)
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-extra-configs --e2e-steps 0 --cpu-budget 1 > $out/bench_under_ncu_${tag}.log 2>&1
# 2. full captures
timeout -s KILL 300 $NCU --set full -k regex:hessian_big_kernel -c 1 -o $out/hess_big_c3_${tag} -f \
    python tools/prof_fused.py 256 1e7 hessian > $out/prof1_${tag}.log 2>&1
timeout -s KILL 300 $NCU --set full -k regex:pass_fused_kernel --launch-skip 1 -c 1 -o $out/fused_wst_c3_${tag} -f \
    python tools/prof_fused.py 256 1e7 hessian > $out/prof2_${tag}.log 2>&1
timeout -s KILL 300 $NCU --set full -k regex:pass_fused_kernel --launch-skip 2 -c 1 -o $out/fused_c3_${tag} -f \
    python tools/prof_fused.py 256 1e7 fused > $out/prof3_${tag}.log 2>&1
timeout -s KILL 300 $NCU --set full --import-source on -k regex:hessian_small_kernel --launch-skip 1 -c 1 -o $out/hess_small_c2_${tag} -f \
    python tools/prof_fused.py 64 1e6 hessian > $out/prof4_${tag}.log 2>&1
timeout -s KILL 300 $NCU --set full -k regex:weights_kernel -c 1 -o $out/weights_c3_${tag} -f \
    python tools/prof_fused.py 256 4e6 moments > $out/prof5_${tag}.log 2>&1
ls -la $out/*${tag}*
