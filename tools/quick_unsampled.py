import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem
K, N = 256, 4*10**6
N_k = np.full(K, N // (K - 3), float); N_k[[5, 100, 255]] = 0; N_k[0] += N - N_k.sum()
p = DeviceProblem(None, N_k, N_local=N)
p.synthesize(np.linspace(1, 5, K), np.linspace(1, 3, K), seed=0)
f = np.zeros(K)
for kern in ("auto", "generic"):
    p.set_kernel(kern)
    p.self_consistent_update(f)
    t = time.time(); out = p.self_consistent_update(f); dt = time.time() - t
    print(kern, f"all-state SCI call {dt*1e3:.2f} ms, pass kernel {p.last_pass_ms():.3f} ms; f'[5,100,255] =", out[[5, 100, 255]])
