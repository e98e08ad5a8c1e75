"""BAR initialisation (mbar.py:1936-1988): pymbar_b200.initialize vs the reference's
`MBAR._initialize_with_bar` where the checkout exists, and vs the identity BAR == two-state MBAR everywhere."""
import os
import sys

import numpy as np
import pytest

from oracle import mbar_oracle as orc
from oracle import testsystems as ots
from pymbar_b200.initialize import bar_delta_f, initialize_with_bar

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bar_equals_two_state_mbar():
    _, u_kn, N_k = ots.harmonic_u_kn([0.0, 1.5], [1.0, 3.0], [400, 250], seed=5)
    f = orc.mbar_f_k(u_kn, N_k)
    w_F = u_kn[1, :400] - u_kn[0, :400]
    w_R = u_kn[0, 400:] - u_kn[1, 400:]
    dF = bar_delta_f(w_F, w_R, rtol=1e-12)
    assert abs(dF - (f[1] - f[0])) < 1e-9


def test_chain_with_an_empty_state_and_a_start_vector():
    _, u_kn, N_k = ots.harmonic_u_kn([0, 1, 2, 3], [1, 2, 4, 8], [100, 50, 0, 80], seed=3)
    x_k = np.repeat(np.arange(4), N_k)
    f = initialize_with_bar(u_kn, N_k, x_k)
    assert f[0] == 0 and f[2] == 0                      # unsampled states are left alone
    pair = np.concatenate([np.arange(0, 100), np.arange(100, 150)])
    f2 = orc.mbar_f_k(u_kn[:2][:, pair], np.array([100, 50]))
    assert abs(f[1] - f2[1]) < 1e-4
    g = initialize_with_bar(u_kn, N_k, x_k, f_k_init=np.array([5.0, 0, 0, 0]))
    s = N_k > 0
    np.testing.assert_allclose((g - g[0])[s], (f - f[0])[s], atol=1e-4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/pymbar"), reason="reference checkout not present")
def test_matches_the_reference_initialisation():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
    sys.path.insert(0, "/root/reference")
    os.environ["PYMBAR_DISABLE_JAX"] = "1"
    try:
        import pymbar
        from pymbar import testsystems

        tc = testsystems.HarmonicOscillatorsTestCase(O_k=[0, 1, 2, 3, 4], K_k=[1, 2, 4, 8, 16])
        _, u, N, _ = tc.sample([300, 200, 0, 250, 100], mode="u_kn", seed=3)
        m = pymbar.MBAR(u, N, initialize="zeros")
        ref = m._initialize_with_bar(m.u_kn)
        mine = initialize_with_bar(m.u_kn, m.N_k, m.x_kindices)
        assert np.max(np.abs(ref - mine)) < 1e-4          # both solve Bennett's equation to rtol 1e-5
    finally:
        sys.path.remove("/root/reference")
        sys.path.remove(os.path.join(ROOT, "oracle", "ref_shim"))
