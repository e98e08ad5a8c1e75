"""Randomised shape sweep (GPU): K in [1, 2100], N in [1, 4000], empty states at random, random f_k and
random per-sample multiplicities — every primitive against the oracle.  Seeds are fixed; the sweep is
deterministic.  Catches tiling / masking / tail-stage edge cases that hand-picked shapes miss."""
import numpy as np
import pytest

from oracle import mbar_oracle as orc

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.RandomState(seed)
    K = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 17, 31, 32, 33, 48, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256,
                        257, 300, 384, 511, 512, 513, 520, 777, 1024, 1100, 2048, 2100]))
    per = int(rng.randint(1, max(2, 4000 // K)))
    N_k = rng.randint(0, per + 1, size=K)
    if rng.rand() < 0.5:
        N_k[N_k == 0] = 1                       # half of the cases: every state sampled
    if N_k.sum() == 0:
        N_k[rng.randint(K)] = 3
    N = int(N_k.sum())
    offs = rng.normal(scale=3.0, size=(K, 1))
    u = rng.normal(scale=2.0, size=(K, N)) + offs
    if rng.rand() < 0.3:
        u[rng.randint(K), rng.randint(N)] = np.inf      # a forbidden configuration
        u[:, N_k.cumsum()[-1] - 1] = np.where(np.isinf(u[:, -1]), 0.0, u[:, -1])
    f = rng.normal(scale=1.0, size=K)
    f -= f[0]
    return K, N, N_k, u, f, rng


@pytest.mark.parametrize("seed", range(48))
def test_random_shape(seed):
    import pymbar_b200

    K, N, N_k, u, f, rng = _case(seed)
    Nf = N_k.astype(float)
    s = Nf > 0
    # a sample must have a finite energy in at least one sampled state
    bad = ~np.isfinite(u[s]).any(axis=0)
    u[np.ix_(np.flatnonzero(s)[:1], np.flatnonzero(bad))] = 0.0
    with pymbar_b200.DeviceProblem(u, Nf) as p:
        np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(u, Nf, f), atol=1e-9)
        np.testing.assert_allclose(p.gradient(f)[s], orc.mbar_gradient(u[s], Nf[s], f[s]), rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(p.objective(f), orc.mbar_objective(u[s], Nf[s], f[s]), rtol=1e-10, atol=1e-8)
        if K <= 130 or seed % 3 == 0:
            np.testing.assert_allclose(p.hessian(f)[np.ix_(s, s)], orc.mbar_hessian(u[s], Nf[s], f[s]),
                                       rtol=1e-8, atol=1e-9)
        if K * N <= 400_000:
            lw = p.log_W_nk(f)
            ref = orc.mbar_log_W_nk(u, Nf, f)
            fin = np.isfinite(ref)
            np.testing.assert_allclose(lw[fin], ref[fin], atol=1e-9)
        if s.sum() > 1:
            f5 = p.sci_iterate(f, 3)
            fh = f.copy()
            for _ in range(3):
                nxt = orc.self_consistent_update(u[s], Nf[s], fh[s])
                fh[s] = nxt - nxt[0]
            np.testing.assert_allclose(f5[s], fh[s], atol=1e-9)
        # per-sample multiplicities == the gathered problem
        w = rng.randint(0, 3, size=N).astype(float)
        for k in np.flatnonzero(s):
            blk = slice(int(N_k[:k].sum()), int(N_k[:k + 1].sum()))
            if w[blk].sum() == 0:
                w[blk.start] = 1.0
        idx = np.repeat(np.arange(N), w.astype(int))
        p.set_sample_weights(w)
        np.testing.assert_allclose(p.gradient(f)[s], orc.mbar_gradient(u[s][:, idx], Nf[s], f[s]), rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(u[:, idx], Nf, f), atol=1e-9)
