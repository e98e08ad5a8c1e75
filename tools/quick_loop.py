import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem
K, N = 256, 10**7
N_k = np.full(K, N // K, float); N_k[-1] += N - N_k.sum()
p = DeviceProblem(None, N_k, N_local=N)
p.synthesize(np.linspace(1, 5, K), np.linspace(1, 3, K), seed=0)
f = np.zeros(K)
p.set_kernel("fused")
for it in (5, 10, 50, 200, 200):
    f = p.sci_iterate(f, it)
    l = p.last_loop_ms()
    print(f"sci_iterate {it}: loop {l['total_ms']/it:.3f} ms/iter, kernel {l['kernel_ms_sum']/it:.3f} ms -> {8*K*N/(l['kernel_ms_sum']/it)/1e6:.0f} GB/s", flush=True)
ms = []
for _ in range(10):
    p.gradient(f); ms.append(p.last_pass_ms())
print("gradient passes:", min(ms), np.median(ms))
os.system("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv")
