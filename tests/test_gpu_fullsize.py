"""Parity at BASELINE.json's full sizes, where the CPU oracle cannot hold the problem: size-independent
properties of the solution (the four invariants of pymbar/tests/test_mbar_solvers.py:34-41), additivity
over sample shards, agreement of a downloaded slice with the oracle, and the analytic free energies of
the harmonic-oscillator family (testsystems/harmonic_oscillators.py:92-96)."""
import numpy as np
import pytest

from oracle import mbar_oracle as orc
from oracle import testsystems as ots

pytestmark = pytest.mark.gpu


def _problem(lib, K, N, seed=0, O=None, k=None):
    N_k = np.full(K, N // K, float)
    N_k[-1] += N - N_k.sum()
    O = np.linspace(1, 5, K) if O is None else O
    k = np.linspace(1, 3, K) if k is None else k
    p = lib.DeviceProblem(None, N_k, N_local=N)
    p.synthesize(O, k, seed=seed)
    return p, N_k, O, k


@pytest.fixture(scope="module")
def lib():
    import pymbar_b200

    return pymbar_b200


@pytest.mark.parametrize("K,N", [(64, 1_000_000), (256, 10_000_000), (32, 10_000_000)])
def test_invariants_at_baseline_sizes(lib, K, N):
    p, N_k, O, kk = _problem(lib, K, N)
    try:
        f, r = p.solve_adaptive(np.zeros(K), tol=1e-12, min_sc_iter=0)
        assert r["success"], r
        # (1) gradient vanishes, (2) sum_n W_nk = 1, (4) self-consistent update is a fixed point
        S, sumL, _ = p.streaming_pass(f)
        np.testing.assert_allclose(S, 1.0, atol=1e-9)
        assert np.max(np.abs(p.gradient(f))) < 1e-8 * N_k.max()
        np.testing.assert_allclose(p.self_consistent_update(f), f, atol=1e-9)
        # (3) sum_k N_k W_nk = 1 on a slice: exp(log W) . N_k, via the per-sample log-denominators
        n0, n = N // 3, 2048
        sl = p.download(n0, n)
        L_ref = orc.log_denominator_n(sl, N_k, f)
        W = np.exp(f[None, :] - sl.T - L_ref[:, None])
        np.testing.assert_allclose(W @ N_k, 1.0, atol=1e-10)
        # slice agreement with the oracle: same per-sample L_n, same partial S_k on those samples
        with lib.DeviceProblem(sl, N_k) as q:
            Sq, sumLq, _ = q.streaming_pass(f)
            S_o, L_o = orc.single_pass_sums(sl, N_k, f)
            np.testing.assert_allclose(Sq, S_o, rtol=1e-11)
            np.testing.assert_allclose(sumLq, L_o.sum(), rtol=1e-12)
        # fused and generic kernels agree at full size
        p.set_kernel("generic")
        Sg, sumLg, _ = p.streaming_pass(f)
        p.set_kernel("auto")
        np.testing.assert_allclose(Sg, S, rtol=1e-11)
        np.testing.assert_allclose(sumLg, sumL, rtol=1e-12)
        # analytic answer of the family, to statistical accuracy
        fa = ots.harmonic_analytical_f_k(kk)
        assert np.max(np.abs(f - fa)) < 0.02, np.max(np.abs(f - fa))
    finally:
        p.close()


def test_shard_additivity_and_k512(lib):
    """C5 shape per GPU (K = 512, two-CTA clusters): partial sums of two sample shards add up to the
    sums over their union (what the per-iteration all-reduce relies on)."""
    K, N = 512, 2_000_000
    N_k = np.full(K, 2 * N // K, float)
    O, kk = np.linspace(1, 5, K), np.linspace(1, 3, K)
    f = np.linspace(0, 1, K)
    parts = []
    for r in range(2):
        p = lib.DeviceProblem(None, N_k, N_local=N)
        p.synthesize(O, kk, seed=3, n_offset=r * N, N_global=2 * N)
        S, sumL, _ = p.streaming_pass(f)
        parts.append((S, sumL))
        p.close()
    p = lib.DeviceProblem(None, N_k, N_local=2 * N)
    p.synthesize(O, kk, seed=3, n_offset=0, N_global=2 * N)
    S, sumL, _ = p.streaming_pass(f)
    np.testing.assert_allclose(parts[0][0] + parts[1][0], S, rtol=1e-12)
    np.testing.assert_allclose(parts[0][1] + parts[1][1], sumL, rtol=1e-13)
    sl = p.download(12345, 512)
    with lib.DeviceProblem(sl, N_k) as q:
        Sq, _, _ = q.streaming_pass(f)
        S_o, _ = orc.single_pass_sums(sl, N_k, f)
        np.testing.assert_allclose(Sq, S_o, rtol=1e-11)
    p.close()


def test_c5_per_gpu_shard_shape(lib):
    """BASELINE.json configs[4] per GPU: K = 512, N = 1.25e7 (51.2 GB of u_kn in HBM).  One device-resident
    self-consistent iteration from the analytic free energies must stay at them to statistical accuracy,
    sum_n W_nk must be ~1 there, and a downloaded slice must agree with the oracle."""
    K, N = 512, 12_500_000
    p, N_k, O, kk = _problem(lib, K, N, seed=1)
    try:
        fa = ots.harmonic_analytical_f_k(kk)
        S, sumL, _ = p.streaming_pass(fa)
        assert np.max(np.abs(S - 1.0)) < 0.02
        f1 = p.sci_iterate(fa, 1)
        assert np.max(np.abs(f1 - fa)) < 0.02
        sl = p.download(N - 4096 - 7, 4096)
        with lib.DeviceProblem(sl, N_k) as q:
            Sq, _, _ = q.streaming_pass(fa)
            S_o, _ = orc.single_pass_sums(sl, N_k, fa)
            np.testing.assert_allclose(Sq, S_o, rtol=1e-11)
        assert p.last_pass_ms() < 60.0
    finally:
        p.close()
