import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import testsystems as ots
from oracle import mbar_oracle as orc
from pymbar_b200 import DeviceProblem

def problem(K, N, seed, empty=()):
    u, N_k = ots.oscillators(K, max(1, N // K), seed=seed)
    N_k = N_k.astype(float)
    for e in empty: N_k[e] = 0.0
    rng = np.random.RandomState(seed); f = rng.normal(scale=0.5, size=K); f -= f[0]
    return u, N_k, f

for K in (48, 40, 192):
    for empty in ((), (3,)):
        u, N_k, f = problem(K, 30 * K, 900 + K, empty)
        for tag, env, pre in (("plain", {}, False), ("hess-first", {}, True), ("nograph", {"MBAR_B200_NO_GRAPH": "1"}, True),
                              ("nom2", {"MBAR_B200_NO_M2": "1"}, True)):
            for k, v in env.items(): os.environ[k] = v
            with DeviceProblem(u, N_k) as p:
                if pre:
                    p.streaming_pass(f); p.hessian(f)
                out = []
                for mode in ("device", "stepped"):
                    p.set_loop_mode(mode)
                    fk, r = p.solve_adaptive(np.zeros(K), tol=1e-12, maxiter=60, min_sc_iter=0)
                    out.append((mode, r["success"], r["iterations"], r["nr_iterations"], r["sci_iterations"], "%.2e" % r["max_delta"], "%.2e" % r["gnorm"]))
                print(K, empty, tag, out, p.loop_stats(), flush=True)
            for k in env: del os.environ[k]
    ref = orc.adaptive(u, N_k, np.zeros(K), tol=1e-12, options=dict(min_sc_iter=0, maxiter=60))
    print("oracle", K, ref["success"], len(ref.get("history", [])))
