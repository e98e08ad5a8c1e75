"""GPU parity: the CUDA path (through the C ABI) vs the CPU oracle and the reference-generated
fixtures, on the same inputs.  Tolerances: |delta f| < 1e-8 (BASELINE.json north_star); primitives
are held much tighter (1e-10 .. 1e-12) because only reduction order and exp rounding differ."""
import numpy as np
import pytest

from oracle import mbar_oracle as orc
from tests import _cases

pytestmark = pytest.mark.gpu

TOL_F = 1e-8


@pytest.fixture(scope="module")
def lib():
    import pymbar_b200
    from pymbar_b200 import _lib

    _lib.load()
    if _lib.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return pymbar_b200


@pytest.mark.parametrize("kernel", ["fused", "generic"])
@pytest.mark.parametrize("name", _cases.ALL)
def test_primitives(lib, name, kernel):
    z = _cases.load(name)
    u, N = z["u_kn"], z["N_k"].astype(float)
    s = N > 0
    K = len(N)
    with lib.DeviceProblem(u, N) as p:
        p.set_kernel(kernel if (kernel == "generic" or K <= 256) else "auto")
        for tag, f in (("zero", np.zeros(K)), ("rand", z["f_rand"])):
            # self-consistent update: all states, fixture produced by the real reference
            np.testing.assert_allclose(p.self_consistent_update(f), z[f"{tag}_sci"], rtol=0, atol=1e-10)
            g = p.gradient(f)
            np.testing.assert_allclose(g[s], z[f"{tag}_grad"], rtol=1e-10, atol=1e-9 * N.max())
            assert np.all(g[~s] == 0)
            # objective: reference value is for the sampled sub-problem; unsampled rows have N_k f_k = 0
            np.testing.assert_allclose(p.objective(f), z[f"{tag}_obj"], rtol=1e-12, atol=1e-7)
            H = p.hessian(f)[np.ix_(s, s)]
            if f"{tag}_hess" in z:
                np.testing.assert_allclose(H, z[f"{tag}_hess"], rtol=1e-10, atol=1e-10)
                np.testing.assert_allclose(p.log_W_nk(f), z[f"{tag}_logW"], rtol=0, atol=1e-10)
            else:
                np.testing.assert_allclose(np.diag(H), z[f"{tag}_hess_diag"], rtol=1e-10)
                np.testing.assert_allclose(np.linalg.norm(H), z[f"{tag}_hess_fro"], rtol=1e-10)
                np.testing.assert_allclose(p.log_W_nk(f)[:4], z[f"{tag}_logW_head"], atol=1e-10)


@pytest.mark.parametrize("name", _cases.SMALL + ["osc_50x100"])
def test_pass_outputs_vs_oracle(lib, name):
    z = _cases.load(name)
    u, N = z["u_kn"], z["N_k"].astype(float)
    s = N > 0
    f = z["f_rand"]
    with lib.DeviceProblem(u, N) as p:
        S, sumL, G = p.streaming_pass(f, want_G=True)
        S_ref, L_ref = orc.single_pass_sums(u[s], N[s], f[s])
        np.testing.assert_allclose(S[s], S_ref, rtol=1e-11)
        np.testing.assert_allclose(sumL, L_ref.sum(), rtol=1e-12)
        W = orc.mbar_W_nk(u[s], N[s], f[s])
        np.testing.assert_allclose(G[np.ix_(s, s)], W.T @ W, rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(p.log_denominator(f), orc.log_denominator_n(u[s], N[s], f[s]), atol=1e-11)
        W_dev = p.log_W_nk(f, exponentiate=True)
        np.testing.assert_allclose(W_dev[:, s], W, rtol=1e-10, atol=1e-300)


@pytest.mark.parametrize("proto", ["default", "robust", "adaptive"])
@pytest.mark.parametrize("name", _cases.ALL)
def test_solved_f_k(lib, name, proto):
    """solve_mbar_for_all_states through the mirrored module API == reference MBAR.f_k."""
    z = _cases.load(name)
    ms = lib.mbar_solvers
    protocol = {"default": ms.DEFAULT_SOLVER_PROTOCOL, "robust": ms.ROBUST_SOLVER_PROTOCOL,
                "adaptive": ms.BOOTSTRAP_SOLVER_PROTOCOL}[proto]
    protocol = tuple({k: (dict(v) if isinstance(v, dict) else v) for k, v in st.items()} for st in protocol)
    N_k = z["N_k"]
    sws = np.where(N_k != 0)[0]
    f = ms.solve_mbar_for_all_states(z["u_kn"], N_k, np.zeros(len(N_k)), sws, protocol)
    assert np.max(np.abs(f - z[f"fk_{proto}"])) < TOL_F
    # the four invariants of pymbar/tests/test_mbar_solvers.py:34-41, evaluated by the ORACLE
    Nf = N_k.astype(float)
    s = Nf > 0
    np.testing.assert_allclose(orc.mbar_gradient(z["u_kn"][s], Nf[s], f[s]), 0, atol=1e-8 * max(1, Nf.max() / 100))
    np.testing.assert_allclose(orc.self_consistent_update(z["u_kn"], Nf, f) - f, 0, atol=1e-10)
    W = orc.mbar_W_nk(z["u_kn"], Nf, f)
    np.testing.assert_allclose(W.sum(0), 1, atol=1e-9)
    np.testing.assert_allclose(W @ Nf, 1, atol=1e-9)


@pytest.mark.parametrize("name", ["small_osc_8x40", "small_empty_state", "osc_50x100"])
def test_native_loops(lib, name):
    z = _cases.load(name)
    u, N = z["u_kn"], z["N_k"].astype(float)
    s = N > 0
    K = len(N)
    with lib.DeviceProblem(u, N) as p:
        f_ad, r = p.solve_adaptive(np.zeros(K), tol=1e-12, min_sc_iter=0)
        assert r["success"]
        ref = np.zeros(K)
        ref[s] = z["adaptive_x"]
        assert np.max(np.abs(f_ad[s] - ref[s])) < TOL_F
        f_sci, r2 = p.solve_sci(np.zeros(K), tol=1e-13, maxiter=20000)
        assert r2["success"] and np.max(np.abs(f_sci[s] - ref[s])) < 1e-7
        # device-resident iteration == host-stepped self-consistent iteration
        f0 = np.zeros(K)
        f_dev = p.sci_iterate(f0, 5)
        f_host = f0.copy()
        for _ in range(5):
            nxt = orc.self_consistent_update(u[s], N[s], f_host[s])
            f_host[s] = nxt - nxt[0]
        np.testing.assert_allclose(f_dev[s], f_host[s], atol=1e-11)


def test_upload_layouts_and_roundtrip(lib):
    z = _cases.load("small_osc_8x40")
    u, N = z["u_kn"], z["N_k"].astype(float)
    f = z["f_rand"]
    with lib.DeviceProblem(u, N) as p:
        want = p.self_consistent_update(f)
        back = p.download()
        np.testing.assert_allclose(back, u, rtol=0, atol=1e-12 * np.abs(u).max())
    big = np.zeros((u.shape[0], u.shape[1] + 7))
    big[:, 3:3 + u.shape[1]] = u
    view = big[:, 3:3 + u.shape[1]]                      # non-contiguous rows (ld > N)
    pin = lib.PinnedArray(u.shape)
    pin.array[:] = u
    for src in (view, pin.array, np.asfortranarray(u), u.astype(np.float32).astype(np.float64)):
        with lib.DeviceProblem(src, N) as p:
            np.testing.assert_allclose(p.self_consistent_update(f), want, atol=1e-12)
    pin.free()


def test_errors(lib):
    from pymbar_b200._lib import MbarB200Error

    z = _cases.load("small_exp_6x50")
    u, N = z["u_kn"].copy(), z["N_k"].astype(float)
    u[2, 5] = np.nan
    with pytest.raises(MbarB200Error) as e:
        lib.DeviceProblem(u, N)
    assert e.value.status == -5
    u[2, 5] = np.inf                                      # +inf energy is legal: weight 0
    with lib.DeviceProblem(u, N) as p:
        f = np.zeros(len(N))
        ref = orc.self_consistent_update(u, N, f)
        np.testing.assert_allclose(p.self_consistent_update(f), ref, atol=1e-10)
        with pytest.raises(MbarB200Error) as e2:          # raw C ABI: explicit range error
            p.streaming_pass(np.full(len(N), 1e9))
        assert e2.value.status == -6
        assert np.all(np.isnan(p.gradient(np.full(len(N), np.nan))))   # mirror: NaN in, NaN out
    with pytest.raises(lib.ParameterError):
        lib.mbar_solvers.solve_mbar_once(z["u_kn"], N, np.zeros(len(N)), method="no-such-method")
    with pytest.raises(TypeError):
        lib.mbar_solvers.validate_inputs(z["u_kn"], list(N), np.zeros(len(N)))


def test_wide_f_spread_falls_back(lib):
    """A spread of f_k + log N_k beyond the fused kernel's range must still give oracle results."""
    z = _cases.load("small_osc_8x40")
    u, N = z["u_kn"], z["N_k"].astype(float)
    f = np.linspace(0, 3000, len(N))
    with lib.DeviceProblem(u, N) as p:
        np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(u, N, f), atol=1e-9)


def test_large_energy_offsets(lib):
    """States whose energies carry large constant offsets (true f_k thousands of kT from the
    starting guess): S_k underflows in linear arithmetic, the reference's logsumexp does not."""
    z = _cases.load("small_osc_8x40")
    N = z["N_k"].astype(float)
    K = len(N)
    off = np.linspace(0, 5000, K)
    u = z["u_kn"] + off[:, None]
    f0 = np.zeros(K)
    with lib.DeviceProblem(u, N) as p:
        np.testing.assert_allclose(p.self_consistent_update(f0), orc.self_consistent_update(u, N, f0), atol=1e-8)
        np.testing.assert_allclose(p.gradient(f0), orc.mbar_gradient(u, N, f0), rtol=1e-9, atol=1e-7)
        f_dev = p.sci_iterate(f0, 3)
        f_host = f0.copy()
        for _ in range(3):
            nxt = orc.self_consistent_update(u, N, f_host)
            f_host = nxt - nxt[0]
        np.testing.assert_allclose(f_dev, f_host, atol=1e-8)
    f = lib.mbar_solvers.solve_mbar_for_all_states(u, z["N_k"], np.zeros(K), np.arange(K),
                                                   lib.mbar_solvers.DEFAULT_SOLVER_PROTOCOL)
    np.testing.assert_allclose(f - off, z["fk_default"], atol=1e-7)


@pytest.mark.parametrize("K,n", [(512, 12), (300, 20), (257, 8), (384, 10), (513, 4), (700, 3), (1024, 3),
                                 (1025, 2), (1500, 2), (2048, 2)])
def test_two_cta_cluster_kernel(lib, K, n):
    """256 < K <= 2048: the fused kernel runs as clusters of 2 / 4 / 8 CTAs (K/CL states each, partial
    denominators exchanged through distributed shared memory).  Checked against the oracle."""
    from oracle import testsystems as ots

    u, N_k = ots.oscillators(K, n, seed=100 + K)
    N = N_k.astype(float)
    rng = np.random.RandomState(K)
    f = rng.normal(scale=0.5, size=K)
    f -= f[0]
    with lib.DeviceProblem(u, N) as p:
        p.set_kernel("fused")
        for mode_env in ("3", "0"):
            import os
            os.environ["MBAR_B200_FUSED_MODE"] = mode_env
            S, sumL, _ = p.streaming_pass(f)
            S_ref, L_ref = orc.single_pass_sums(u, N, f)
            np.testing.assert_allclose(S, S_ref, rtol=1e-11)
            np.testing.assert_allclose(sumL, L_ref.sum(), rtol=1e-12)
        os.environ.pop("MBAR_B200_FUSED_MODE", None)
        np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(u, N, f), atol=1e-10)
        f5 = p.sci_iterate(np.zeros(K), 4)
        fh = np.zeros(K)
        for _ in range(4):
            nxt = orc.self_consistent_update(u, N, fh)
            fh = nxt - nxt[0]
        np.testing.assert_allclose(f5, fh, atol=1e-10)
        H = p.hessian(f)
        np.testing.assert_allclose(H, orc.mbar_hessian(u, N, f), rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("K,N_k", [(1, [5]), (2, [1, 2]), (3, [1, 0, 1]), (2, [40, 0]), (7, [3] * 7), (33, [2] * 33),
                                   (5, [31, 1, 0, 0, 1]), (4, [32, 32, 32, 32])])
def test_tiny_and_ragged_shapes(lib, K, N_k):
    """Degenerate sizes: a single state, fewer than 32 samples (one partial tile), exactly one tile,
    N_k = 0 in various places.  Every primitive against the oracle."""
    rng = np.random.RandomState(K * 100 + sum(N_k))
    N_k = np.array(N_k)
    N = int(N_k.sum())
    u = rng.normal(size=(K, N)) * 3 + rng.normal(size=(K, 1)) * 5
    Nf = N_k.astype(float)
    s = Nf > 0
    f = rng.normal(size=K)
    f -= f[0]
    for kern in ("fused", "generic"):
        with lib.DeviceProblem(u, Nf) as p:
            p.set_kernel(kern)
            np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(u, Nf, f), atol=1e-10)
            np.testing.assert_allclose(p.gradient(f)[s], orc.mbar_gradient(u[s], Nf[s], f[s]), rtol=1e-10, atol=1e-9)
            np.testing.assert_allclose(p.objective(f), orc.mbar_objective(u[s], Nf[s], f[s]), rtol=1e-11, atol=1e-9)
            np.testing.assert_allclose(p.hessian(f)[np.ix_(s, s)], orc.mbar_hessian(u[s], Nf[s], f[s]), rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(p.log_W_nk(f), orc.mbar_log_W_nk(u, Nf, f), atol=1e-10)
    sws = np.where(N_k != 0)[0]
    proto = tuple(dict(st) for st in lib.mbar_solvers.DEFAULT_SOLVER_PROTOCOL)
    got = lib.mbar_solvers.solve_mbar_for_all_states(u, N_k, np.zeros(K), sws, proto)
    want = orc.solve_mbar_for_all_states(u, N_k, np.zeros(K), sws, orc.DEFAULT_SOLVER_PROTOCOL)
    np.testing.assert_allclose(got, want, atol=1e-8)


def test_cluster_kernel_with_unsampled_states(lib):
    """K > 256 (two-CTA clusters) together with empty states (weight e^-80 ride-along) and ragged N."""
    from oracle import testsystems as ots

    K = 300
    N_k = np.full(K, 7)
    N_k[[0, 17, 150, 299]] = 0
    O, kk = np.linspace(1, 5, K), np.linspace(1, 3, K)
    _, u, _ = ots.harmonic_u_kn(O, kk, N_k, seed=5)
    Nf = N_k.astype(float)
    f = np.random.RandomState(1).normal(scale=0.3, size=K)
    with lib.DeviceProblem(u, Nf) as p:
        for kern in ("auto", "generic"):
            p.set_kernel(kern)
            np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(u, Nf, f), atol=1e-10)
        p.set_kernel("auto")
        S, G = p.weight_moments(f)
        W = orc.mbar_W_nk(u, Nf, f)
        np.testing.assert_allclose(S, W.sum(0), rtol=1e-10)
        np.testing.assert_allclose(G, W.T @ W, rtol=1e-9, atol=1e-14)


def test_one_shot_host_entry(lib):
    """mbar_b200_self_consistent_update_host: create + upload + pass + destroy in one C call."""
    import ctypes as C

    from pymbar_b200 import _lib

    z = _cases.load("small_empty_state")
    u = np.ascontiguousarray(z["u_kn"])
    N = z["N_k"].astype(float)
    f = z["f_rand"].copy()
    out = np.empty_like(f)
    dp = C.POINTER(C.c_double)
    rc = _lib.load().mbar_b200_self_consistent_update_host(0, u.shape[0], u.shape[1], C.c_void_p(u.ctypes.data),
                                                           u.shape[1], N.ctypes.data_as(dp), f.ctypes.data_as(dp),
                                                           out.ctypes.data_as(dp))
    assert rc == 0
    np.testing.assert_allclose(out, z["rand_sci"], atol=1e-10)


def test_context_housekeeping(lib):
    """get_shape, counters, the parked-buffer pool and trim()."""
    import ctypes as C

    from pymbar_b200 import _lib

    z = _cases.load("small_osc_8x40")
    N = z["N_k"].astype(float)
    p = lib.DeviceProblem(z["u_kn"], N)
    K, Nl = C.c_int32(0), C.c_int64(0)
    assert _lib.load().mbar_b200_get_shape(p._h, C.byref(K), C.byref(Nl)) == 0
    assert (K.value, Nl.value) == z["u_kn"].shape
    before = p.counters()
    p.gradient(np.zeros(len(N)))
    after = p.counters()
    assert after["passes"] == before["passes"] + 1 and after["launches"] > before["launches"]
    want = p.self_consistent_update(z["f_rand"])
    p.close()
    q = lib.DeviceProblem(z["u_kn"], N)            # reuses the parked buffers
    np.testing.assert_array_equal(q.self_consistent_update(z["f_rand"]), want)
    q.close()
    lib.trim()
    r = lib.DeviceProblem(z["u_kn"], N)            # fresh allocations after trim
    np.testing.assert_array_equal(r.self_consistent_update(z["f_rand"]), want)
    r.close()
