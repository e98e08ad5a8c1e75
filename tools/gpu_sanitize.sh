#!/bin/bash
# compute-sanitizer over the round-2 kernels on small parity cases:  gpurun -- 'bash tools/gpu_sanitize.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
SEL="test_hessian_kernel_paths_vs_oracle or test_candidate_batched_pass_kernel or test_24_row_kernel_family or test_adaptive_iteration_as_cuda_graph or test_pass_multi or test_device_newton_step_sizes"
timeout -s KILL 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $out/memcheck_${tag}.log \
    python -m pytest tests/test_gpu_loops.py -q -m gpu -p no:cacheprovider -x -k "$SEL" > $out/memcheck_${tag}_pytest.log 2>&1
echo "memcheck exit $?"; tail -3 $out/memcheck_${tag}_pytest.log; grep -E "ERROR SUMMARY|Invalid|Error" $out/memcheck_${tag}.log | head -10
SEL2="test_hessian_kernel_paths_vs_oracle[48] or test_hessian_kernel_paths_vs_oracle[200] or test_candidate_batched_pass_kernel[64] or test_candidate_batched_pass_kernel[700] or test_24_row_kernel_family[96] or test_24_row_kernel_family[768] or test_device_newton_step_sizes[40] or test_device_newton_step_sizes[200]"
timeout -s KILL 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $out/racecheck_${tag}.log \
    python -m pytest tests/test_gpu_loops.py -q -m gpu -p no:cacheprovider -x -k "$SEL2" > $out/racecheck_${tag}_pytest.log 2>&1
echo "racecheck exit $?"; tail -3 $out/racecheck_${tag}_pytest.log; grep -E "RACECHECK SUMMARY|hazard|Race" $out/racecheck_${tag}.log | head -10
