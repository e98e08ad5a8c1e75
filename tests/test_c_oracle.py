"""The C (pthreads) restatement (oracle/mbar_oracle.c) against the numpy oracle and the reference fixtures."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import mbar_oracle as orc
from tests import _cases


@pytest.mark.parametrize("name", _cases.SMALL + ["osc_50x100", "exp_200x50"])
def test_c_oracle_matches_reference(name):
    z = _cases.load(name)
    u, N = z["u_kn"], z["N_k"].astype(float)
    s = N > 0
    for tag, f in (("zero", np.zeros(len(N))), ("rand", z["f_rand"])):
        np.testing.assert_allclose(c_oracle.self_consistent_update(u, N, f), z[f"{tag}_sci"], atol=1e-11)
        np.testing.assert_allclose(c_oracle.mbar_gradient(u, N, f)[s], z[f"{tag}_grad"], rtol=1e-10, atol=1e-9)
    assert c_oracle.threads() >= 1
