"""Pin the CPU oracle against the reference's golden vector and reference-generated fixtures.

CPU only.  The fixtures in tests/golden were produced by oracle/make_golden.py running the
unmodified reference (pymbar @ cf12100f) in the build container.
"""
import numpy as np
import pytest

from oracle import mbar_oracle as orc
from oracle import testsystems as ots
from tests import _cases


def test_golden_example_printed_vector():
    # examples/harmonic-oscillators/harmonic-oscillators.py_output.txt:34-36
    g = ots.GOLDEN_EXAMPLE
    _, u_kn, N_k = ots.harmonic_u_kn(g["O_k"], g["K_k"], g["N_k"], seed=g["seed"])
    f = orc.mbar_f_k(u_kn, N_k)
    assert np.max(np.abs(f - np.array(g["f_k_printed"]))) < 5e-9


@pytest.mark.parametrize("name", _cases.ALL)
def test_primitives_match_reference(name):
    z = _cases.load(name)
    u, N = z["u_kn"], z["N_k"].astype(float)
    s = N > 0
    for tag, f in (("zero", np.zeros(len(N))), ("rand", z["f_rand"])):
        np.testing.assert_allclose(orc.self_consistent_update(u, N, f), z[f"{tag}_sci"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(orc.mbar_gradient(u[s], N[s], f[s]), z[f"{tag}_grad"], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(orc.mbar_objective(u[s], N[s], f[s]), z[f"{tag}_obj"], rtol=1e-13)
        H = orc.mbar_hessian(u[s], N[s], f[s])
        if f"{tag}_hess" in z:
            np.testing.assert_allclose(H, z[f"{tag}_hess"], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(orc.mbar_log_W_nk(u, N, f), z[f"{tag}_logW"], rtol=0, atol=1e-11)
        else:
            np.testing.assert_allclose(np.diag(H), z[f"{tag}_hess_diag"], rtol=1e-12)
            np.testing.assert_allclose(np.linalg.norm(H), z[f"{tag}_hess_fro"], rtol=1e-12)
            np.testing.assert_allclose(orc.mbar_log_W_nk(u, N, f)[:4], z[f"{tag}_logW_head"], atol=1e-11)


@pytest.mark.parametrize("name", _cases.SMALL + ["golden_example", "osc_50x100"])
@pytest.mark.parametrize("proto", ["default", "robust", "adaptive"])
def test_solved_f_k_match_reference(name, proto):
    z = _cases.load(name)
    protocol = {"default": orc.DEFAULT_SOLVER_PROTOCOL, "robust": orc.ROBUST_SOLVER_PROTOCOL,
                "adaptive": orc.BOOTSTRAP_SOLVER_PROTOCOL}[proto]
    f = orc.mbar_f_k(z["u_kn"], z["N_k"], solver_protocol=protocol)
    assert np.max(np.abs(f - z[f"fk_{proto}"])) < 1e-10


@pytest.mark.parametrize("name", _cases.SMALL)
def test_single_pass_formulation(name):
    """The one-exp-per-entry algebra (what the kernels compute) equals the two-logsumexp arithmetic."""
    z = _cases.load(name)
    N = z["N_k"].astype(float)
    s = N > 0
    u, N, f = z["u_kn"][s], N[s], z["f_rand"][s]
    S, L = orc.single_pass_sums(u, N, f)
    np.testing.assert_allclose(f - np.log(S), orc.self_consistent_update(u, N, f), atol=1e-12)
    np.testing.assert_allclose(N * (S - 1), orc.mbar_gradient(u, N, f), rtol=1e-11, atol=1e-10)
    np.testing.assert_allclose(L.sum() - N @ f, orc.mbar_objective(u, N, f), rtol=1e-13)


def test_c1_precondition_matches_reference():
    """precondition_u_kn (mbar_solvers.py:710-735) of the oracle vs probes of the reference's
    output on BASELINE config C1; the C1 solve itself is part of the parametrised tests above."""
    z = _cases.load("c1_harmonic_5x1000")
    u, N = z["u_kn"], z["N_k"].astype(float)
    for tag, f in (("zero", np.zeros(5)), ("rand", z["f_rand"])):
        pc = orc.precondition_u_kn(u, N, f)
        np.testing.assert_allclose(pc[:, :64], z[f"{tag}_precond_head"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(pc.sum(1), z[f"{tag}_precond_rowsum"], rtol=1e-13)
        np.testing.assert_allclose(orc.mbar_objective(pc, N, f), z[f"{tag}_precond_obj"], atol=1e-8)
