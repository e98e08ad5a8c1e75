import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
what = sys.argv[3] if len(sys.argv) > 3 else "fused"
N_k = np.full(K, N // K, float); N_k[-1] += N - N_k.sum()
p = DeviceProblem(None, N_k, N_local=N)
p.synthesize(np.linspace(1, 5, K), np.linspace(1, 3, K), seed=0)
f = np.zeros(K)
if what == "hessian":
    p.hessian(f); p.hessian(f)
    print(p.last_kernels(), p.last_hessian_ms())
elif what == "moments":       # all-state second moments: the weights_kernel path
    p.weight_moments(f)
    print(p.last_kernels(), p.last_hessian_ms())
else:
    p.set_kernel(what)
    for _ in range(3):
        p.gradient(f)
    print(p.last_pass_ms())
