"""Theta / dDelta_f / overlap / N_eff from G = W^T W (K x K) alone, against
outputs of the unmodified reference (MBAR.compute_free_energy_differences, compute_overlap,
compute_effective_sample_number) stored in tests/golden by oracle/make_golden.py."""
import numpy as np
import pytest

from tests import _cases

CASES = _cases.SMALL + ["golden_example", "osc_50x100"]


@pytest.mark.parametrize("name", CASES)
def test_estimators_from_reference_G(name):
    """CPU: the K x K algebra alone, fed with the reference's own W^T W."""
    from pymbar_b200 import estimators as est

    z = _cases.load(name)
    G, N_k = z["est_G"], z["N_k"].astype(float)
    Theta = est.asymptotic_covariance(G, N_k)
    np.testing.assert_allclose(Theta, z["est_Theta"], rtol=1e-7, atol=1e-12)
    r = est.free_energy_differences(z["fk_default"], G, N_k)
    np.testing.assert_allclose(r["Delta_f"], z["est_Delta_f"], atol=1e-12)
    np.testing.assert_allclose(r["dDelta_f"], z["est_dDelta_f"], rtol=1e-6, atol=1e-9)
    ov = est.overlap(G, N_k)
    np.testing.assert_allclose(ov["matrix"], z["est_overlap_matrix"], rtol=1e-12)
    np.testing.assert_allclose(np.real(ov["scalar"]), z["est_overlap_scalar"], atol=1e-10)
    np.testing.assert_allclose(est.effective_sample_number(G), z["est_N_eff"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_estimators_from_device_moments(name):
    """GPU: G for ALL states (sampled or not) from the Hessian pass; no N x K array anywhere."""
    import pymbar_b200
    from pymbar_b200 import estimators as est

    z = _cases.load(name)
    N_k = z["N_k"].astype(float)
    f = z["fk_default"]
    with pymbar_b200.DeviceProblem(z["u_kn"], N_k) as p:
        S, G = p.weight_moments(f)
    np.testing.assert_allclose(S, 1.0, atol=1e-8)                      # tests/test_mbar_solvers.py:37
    np.testing.assert_allclose(G, z["est_G"], rtol=1e-7, atol=1e-13)
    r = est.free_energy_differences(f, G, N_k)
    np.testing.assert_allclose(r["dDelta_f"], z["est_dDelta_f"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(est.effective_sample_number(G), z["est_N_eff"], rtol=1e-7)
    np.testing.assert_allclose(np.real(est.overlap(G, N_k)["scalar"]), z["est_overlap_scalar"], atol=1e-8)
