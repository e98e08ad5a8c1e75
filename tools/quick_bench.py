import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem
cfgs = [(256, 10**7), (64, 10**6), (32, 10**7), (128, 4*10**6)]
if len(sys.argv) > 1:
    cfgs = [(int(a.split("x")[0]), int(float(a.split("x")[1]))) for a in sys.argv[1:]]
for K, N in cfgs:
    N_k = np.full(K, N // K, float); N_k[-1] += N - N_k.sum()
    p = DeviceProblem(None, N_k, N_local=N)
    p.synthesize(np.linspace(1, 5, K), np.linspace(1, 3, K), seed=0)
    f = np.zeros(K)
    p.set_kernel("fused")
    p.gradient(f); ms = []
    for _ in range(10):
        p.gradient(f); ms.append(p.last_pass_ms())
    print(f"[mode {os.environ.get("MBAR_B200_FUSED_MODE","default")}] K={K} N={N:.0e} fused min {min(ms):.3f} med {np.median(ms):.3f} ms -> {8*K*N/min(ms)/1e6:.0f} GB/s", flush=True)
    p.close()
