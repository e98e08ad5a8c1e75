"""expectations and perturbed free energies through the augmented problem
(appended columns = unsampled states), against the unmodified reference's compute_expectations /
compute_perturbed_free_energies at its default solution (fixtures from oracle/make_golden.py)."""
import numpy as np
import pytest

from tests import _cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", _cases.SMALL)
def test_expectations_match_reference(name):
    from pymbar_b200 import expectations as ex

    z = _cases.load(name)
    u, N_k, f, x = z["u_kn"], z["N_k"], z["fk_default"], z["x_n"]
    r = ex.compute_expectations(u, N_k, f, x)
    np.testing.assert_allclose(r["mu"], z["expt_avg_mu"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["sigma"], z["expt_avg_sigma"], rtol=1e-5, atol=1e-8)
    r = ex.compute_expectations(u, N_k, f, x, output="differences")
    np.testing.assert_allclose(r["mu"], z["expt_diff_mu"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["sigma"], z["expt_diff_sigma"], rtol=1e-5, atol=1e-8)
    r = ex.compute_expectations(u, N_k, f, u, state_dependent=True)
    np.testing.assert_allclose(r["mu"], z["expt_sd_mu"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["sigma"], z["expt_sd_sigma"], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("name", _cases.SMALL)
def test_perturbed_free_energies_match_reference(name):
    from pymbar_b200 import expectations as ex

    z = _cases.load(name)
    r = ex.compute_perturbed_free_energies(z["u_kn"], z["N_k"], z["fk_default"], z["pert_u_ln"])
    np.testing.assert_allclose(r["Delta_f"], z["pert_Delta_f"], atol=1e-8)
    np.testing.assert_allclose(r["dDelta_f"], z["pert_dDelta_f"], rtol=1e-5, atol=1e-8)
