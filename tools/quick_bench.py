"""Development timing probe (not the driver's bench): fused/generic pass, Hessian, solves."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem

def run(K, N, passes=10, hess=True, generic=True):
    N_k = np.full(K, N // K, float); N_k[-1] += N - N_k.sum()
    O, kk = np.linspace(1, 5, K), np.linspace(1, 3, K)
    p = DeviceProblem(None, N_k, N_local=N)
    t = time.time(); p.synthesize(O, kk, seed=0); t_s = time.time() - t
    f = np.zeros(K)
    out = {}
    for kern in (["fused"] if K <= 256 else []) + (["generic"] if generic else []):
        p.set_kernel(kern)
        p.gradient(f)
        ms = []
        for _ in range(passes if kern == "fused" else 2):
            p.gradient(f); ms.append(p.last_pass_ms())
        out[kern] = (min(ms), np.median(ms))
    p.set_kernel("auto")
    line = f"K={K} N={N:.0e} synth={t_s:.2f}s " + " ".join(
        f"{k}: min {v[0]:.3f} ms med {v[1]:.3f} ms = {8*K*N/v[0]/1e6:.0f} GB/s" for k, v in out.items())
    if K <= 256:
        t = time.time(); p.sci_iterate(f, 20); dt = (time.time() - t) / 20
        line += f" | sci_iterate {dt*1e3:.3f} ms/iter"
    if hess:
        t = time.time(); p.hessian(f); th = time.time() - t
        line += f" | hessian call {th*1e3:.1f} ms"
        t = time.time(); fk, r = p.solve_adaptive(f, tol=1e-12, min_sc_iter=0); ta = time.time() - t
        line += f" | adaptive {r['iterations']} it ({r['nr_iterations']} nr) {ta:.2f}s gnorm {r['gnorm']:.2e} ok={r['success']}"
    print(line, flush=True)
    p.close()

if __name__ == "__main__":
    for K, N in ((64, 10**6), (32, 10**7), (5, 10**6), (256, 10**6), (256, 10**7)):
        run(K, N)
