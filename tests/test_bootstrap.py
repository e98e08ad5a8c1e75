"""bootstrap replicates as per-sample multiplicities on the resident problem,
against f_k_boots / bootstrap_rints of the unmodified reference (MBAR(n_bootstraps=4, rseed=11))."""
import numpy as np
import pytest

from oracle import mbar_oracle as orc
from tests import _cases


@pytest.mark.parametrize("name", _cases.SMALL)
def test_indices_reproduce_reference_stream(name):
    from pymbar_b200.bootstrap import bootstrap_indices

    z = _cases.load(name)
    got = bootstrap_indices(z["N_k"], 4, 11)
    np.testing.assert_array_equal(got, z["boot_rints"])


@pytest.mark.parametrize("name", _cases.SMALL)
def test_oracle_gather_matches_reference(name):
    """The oracle on the gathered array (what the reference does) reproduces the fixture."""
    z = _cases.load(name)
    N_k = z["N_k"]
    sws = np.where(N_k != 0)[0]
    for b in range(2):
        f = orc.solve_mbar_for_all_states(z["u_kn"][:, z["boot_rints"][b]], N_k, z["boot_f_k_start"].copy(), sws,
                                          orc.BOOTSTRAP_SOLVER_PROTOCOL)
        assert np.max(np.abs(f - z["boot_f_k"][b])) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", _cases.SMALL)
def test_weighted_replicates_match_reference(name):
    import pymbar_b200
    from pymbar_b200.bootstrap import bootstrap_f_k

    z = _cases.load(name)
    N_k = z["N_k"]
    Nf = N_k.astype(float)
    s = Nf > 0
    with pymbar_b200.DeviceProblem(z["u_kn"], Nf) as p:
        boots = bootstrap_f_k(p, z["boot_f_k_start"], N_k, rints=z["boot_rints"])
        assert np.max(np.abs(boots - z["boot_f_k"])) < 1e-8
        again = bootstrap_f_k(p, z["boot_f_k_start"], N_k, n_bootstraps=4, rseed=11)
        np.testing.assert_array_equal(again, boots)                       # determinism under rseed (test_mbar.py:533-545)
        # weighted primitives == primitives on the gathered array (oracle), incl. Hessian and objective
        r = z["boot_rints"][0]
        w = np.bincount(r, minlength=len(r)).astype(float)
        ug = z["u_kn"][:, r]
        f = z["f_rand"]
        for kern in ("fused", "generic"):
            p.set_kernel(kern)
            p.set_sample_weights(w)
            np.testing.assert_allclose(p.self_consistent_update(f), orc.self_consistent_update(ug, Nf, f), atol=1e-10)
            np.testing.assert_allclose(p.gradient(f)[s], orc.mbar_gradient(ug[s], Nf[s], f[s]), rtol=1e-10, atol=1e-9)
            np.testing.assert_allclose(p.objective(f), orc.mbar_objective(ug[s], Nf[s], f[s]), rtol=1e-11, atol=1e-8)
            np.testing.assert_allclose(p.hessian(f)[np.ix_(s, s)], orc.mbar_hessian(ug[s], Nf[s], f[s]), rtol=1e-9, atol=1e-10)
            p.set_sample_weights(None)
            np.testing.assert_allclose(p.gradient(f)[s], orc.mbar_gradient(z["u_kn"][s], Nf[s], f[s]), rtol=1e-10, atol=1e-9)
