import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem
for K, N in ((256, 2*10**6), (256, 10**7), (128, 4*10**6), (64, 10**6), (200, 10**6)):
    N_k = np.full(K, N // K, float); N_k[-1] += N - N_k.sum()
    p = DeviceProblem(None, N_k, N_local=N)
    p.synthesize(np.linspace(1, 5, K), np.linspace(1, 3, K), seed=0)
    f = np.zeros(K)
    p.gradient(f); p.hessian(f)
    t = time.time(); p.gradient(f); tg = time.time() - t
    t = time.time(); H = p.hessian(f); th = time.time() - t
    t = time.time(); fk, r = p.solve_adaptive(f, tol=1e-12, min_sc_iter=0); ta = time.time() - t
    flops = 2.0 * K * K * N
    print(f"K={K} N={N:.0e} grad call {tg*1e3:.2f} ms, hessian call {th*1e3:.2f} ms (hess-only ~{(th-tg)*1e3:.2f} ms = {flops/(th-tg)/1e12:.1f} TFLOP/s full-matrix-equiv) | adaptive {r['iterations']} it {ta:.2f}s ok={r['success']} |H-H^T|={np.abs(H-H.T).max():.1e}", flush=True)
    p.close()
