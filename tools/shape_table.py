"""Fused-pass / generic-pass throughput over a sweep of K (CUDA events of the library, burst: 10 launches after
3 warm-ups) -> markdown table for profiles/.   python tools/shape_table.py > gpurun_out/shapes_r2.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem
import pymbar_b200

PEAK = 6571.9
print("| K | N | kernel | ms / launch | GB/s | frac of copy peak |\n|---|---|---|---:|---:|---:|")
for K, N in ((5, 8_000_000), (32, 10_000_000), (64, 4_000_000), (96, 4_000_000), (128, 4_000_000), (192, 4_000_000),
             (200, 4_000_000), (256, 4_000_000), (384, 3_000_000), (512, 2_000_000), (768, 1_500_000),
             (1024, 1_000_000), (1536, 700_000), (2048, 500_000)):
    N_k = np.full(K, N // K, float); N_k[-1] += N - N_k.sum()
    p = DeviceProblem(None, N_k, N_local=N)
    p.synthesize(np.linspace(1, 5, K), np.linspace(1, 3, K), seed=0)
    f = np.zeros(K)
    p.sci_iterate(f, 3)
    p.sci_iterate(f, 10)
    ms = p.last_loop_ms()["kernel_ms_sum"] / 10
    name = p.last_kernels()["pass_kernel"].split(">")[0].replace("pass_fused_kernel<", "fused<") + ">"
    gbs = 8.0 * K * N / (ms * 1e-3) / 1e9
    print(f"| {K} | {N:.1e} | `{name}` | {ms:.3f} | {gbs:.0f} | {gbs / PEAK:.2f} |", flush=True)
    if K in (256, 2048):
        p.set_kernel("generic")
        for _ in range(2):
            p.gradient(f)
        ms = p.last_pass_ms()
        gbs = 8.0 * K * N / (ms * 1e-3) / 1e9
        print(f"| {K} | {N:.1e} | `pass_generic_kernel` | {ms:.3f} | {gbs:.0f} | {gbs / PEAK:.2f} |", flush=True)
    p.close()
    pymbar_b200.trim()
