import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymbar_b200 import DeviceProblem, PinnedArray
K, N = 256, 4 * 10**6          # 8.2 GB
N_k = np.full(K, N // K, float)
rng = np.random.default_rng(0)
u = rng.random((K, N))
f = np.zeros(K)
for name, arr in (("pageable", u),):
    for _ in range(2):
        t = time.time(); p = DeviceProblem(arr, N_k); t1 = time.time() - t
        t = time.time(); p.self_consistent_update(f); t2 = time.time() - t
        t = time.time(); p.close(); t3 = time.time() - t
        print(f"{name}: create+upload {t1:.3f} s = {arr.nbytes/t1/1e9:.1f} GB/s, pass {t2*1e3:.1f} ms, destroy {t3*1e3:.1f} ms", flush=True)
pin = PinnedArray((K, N)); pin.array[:] = u
for _ in range(2):
    t = time.time(); p = DeviceProblem(pin.array, N_k); t1 = time.time() - t
    p.close()
    print(f"pinned: create+upload {t1:.3f} s = {u.nbytes/t1/1e9:.1f} GB/s", flush=True)
t = time.time(); p = DeviceProblem(None, N_k, N_local=N); print("create only", time.time() - t); p.close()
