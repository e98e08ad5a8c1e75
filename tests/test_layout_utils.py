"""Input-side helpers: vectorised kln_to_kn / kn_to_n == the reference's per-column loops
(pymbar/utils.py:41-114), restated here as the loops they are."""
import numpy as np

from pymbar_b200.utils import kln_to_kn, kn_to_n


def _ref_kln_to_kn(kln, N_k=None):
    K, L, N_max = np.shape(kln)
    if N_k is None:
        N_k = N_max * np.ones([L], dtype=np.int64)
    kn = np.zeros([L, np.sum(N_k)], dtype=np.float64)
    i = 0
    for k in range(K):
        for ik in range(N_k[k]):
            kn[:, i] = kln[k, :, ik]
            i += 1
    return kn


def _ref_kn_to_n(kn, N_k=None):
    K, N_max = np.shape(kn)
    if N_k is None:
        N_k = N_max * np.ones([K], dtype=np.int64)
    n = np.zeros([np.sum(N_k)], dtype=np.float64)
    i = 0
    for k in range(K):
        for ik in range(N_k[k]):
            n[i] = kn[k, ik]
            i += 1
    return n


def test_layout_converters_match_reference_loops():
    rng = np.random.RandomState(0)
    for K, N_max, N_k in ((4, 7, [7, 3, 0, 5]), (3, 5, None), (1, 4, [2]), (5, 6, [0, 0, 6, 1, 0])):
        kln = rng.normal(size=(K, K, N_max))
        Nk = None if N_k is None else np.array(N_k)
        np.testing.assert_array_equal(kln_to_kn(kln, Nk), _ref_kln_to_kn(kln, Nk))
        kn = rng.normal(size=(K, N_max))
        np.testing.assert_array_equal(kn_to_n(kn, Nk), _ref_kn_to_n(kn, Nk))
